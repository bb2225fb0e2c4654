// trace_kernel.cuh — the per-ray device code of libaicb200: a from-scratch sm_100a
// implementation of all-is-cubes' SpaceRaytracer::trace_ray (sr.rs:135-238) and its pixel
// dispatch (renderer.rs:424-451, 516-556).
//
// Design (B200-first, not a translation):
//  * one warp = one 8x4 pixel tile, persistent warps pull tiles from a global atomic counter;
//  * the two-level grid (Space cubes -> block id; block -> N^3 brick of palette indices) is
//    walked by ONE unified Amanatides–Woo DDA whose state lives in registers; entering a
//    recursive block pushes the outer state and re-initialises the same DDA on the brick, so
//    all lanes execute the same step code whatever level they are on;
//  * cell words carry the classification (invisible / single / recursive, voxel invisible) in
//    their top bits, so an empty step costs exactly one dependent 2-byte load;
//  * all ray geometry is f64 and all colour is f32, operation for operation as the reference
//    (compiled with -fmad=false: Rust never contracts to FMA); transcendentals are evaluated in
//    f64 and rounded once.
//
// Every function cites the reference lines it reproduces.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/aicb200.h"

namespace aicb {

// ---- device-side scene -------------------------------------------------------------------------
constexpr uint32_t KIND_INVISIBLE = 0;  // AIR, or a single voxel that is fully transparent + non-emissive
constexpr uint32_t KIND_SINGLE = 1;     // Evoxels::One, visible
constexpr uint32_t KIND_RECURSIVE = 2;  // paletted brick

// 32-byte block record, read as two uint4.
struct BlockRec {
    uint32_t kind_res;     // kind | resolution << 8
    int16_t vlo[3];        // voxel_bounds lower
    uint16_t vsize[3];     // voxel_bounds size
    uint32_t brick_off;    // first u16 of this block's brick in the pool
    uint32_t pal_off;      // first palette entry (single: the voxel)
    uint32_t _pad[2];
};
static_assert(sizeof(BlockRec) == 32, "BlockRec must be 32 bytes");

struct DeviceScene {
    int32_t lo[3];
    int32_t size[3];
    const void *cells;          // u16 (id | kind<<14) or u32 (id | kind<<16), Z-major
    const uint32_t *light;      // PackedLight texels r|g<<8|b<<16|status<<24, or nullptr (== ONE)
    const BlockRec *blocks;
    const uint16_t *bricks;     // palette index | invisible<<15
    const float4 *palette;      // 2 x float4 per entry: rgba, emission
    const float *lut;           // 256-entry PackedLight decode table (data.rs:301-354)
    uint32_t sky_faces[6];      // BlockSky faces NX..PZ as texels (sky.rs:54-82)
    uint32_t sky_mean;
    uint32_t sky_kind;
    float sky_colors[8][3];
    uint32_t wide_cells;        // 0: u16 cells, 1: u32 cells
};

struct TraceParams {
    DeviceScene scene;
    // camera
    double m[16];               // inverse_projection_view, row-major m11..m44
    uint32_t fb_width, fb_height;
    float exposure;
    // options
    uint32_t fog;
    uint32_t lighting;
    uint32_t transparency;
    float threshold;
    uint32_t antialias;
    uint32_t tone_mapping;
    float maximum_intensity;
    double view_distance;
    uint32_t debug_pixel_cost;
    uint32_t include_sky;
    // work description
    uint32_t local_rows;        // rows rendered by this shard
    uint32_t strip_rows, shard_index, shard_count;
    const double *rays;         // explicit rays (trace_rays) or nullptr (camera rays)
    uint64_t n_rays;            // number of explicit rays
    uint32_t tiles_x, tiles_y;
    // outputs
    uchar4 *out_srgb8;
    float4 *out_colorbuf;
    double *out_depth;
    aicb_hit *out_hit;
    uint32_t *out_steps;
    unsigned long long *counters;  // [0] cubes_traced, [1] outer steps, [2] inner steps, [3] hits, [4] light texels, [5] blocks entered
    unsigned int *tile_counter;
};

#ifdef __CUDACC__

#define AICB_DEV __device__ __forceinline__

constexpr int LC_NONE = 0, LC_FLAT = 1, LC_INTERP = 2;  // lighting class (template)

struct Ray {
    double ox, oy, oz, dx, dy, dz;   // original ray (camera space == world space)
    double tdx, tdy, tdz;            // t_delta = 1/|d| (raycast.rs:769)
    int sx, sy, sz;                  // signum_101(d) (raycast.rs:768)
    double half_over_len;            // 0.5 / |d| (raycast.rs:669)
    bool steppable;                  // step != 0 (first clause of valid_for_stepping, raycast.rs:565)
};

// State::* of the active raycaster (raycast.rs:99-121), bounds kept as the per-axis exit limit.
struct Caster {
    double tmx, tmy, tmz;
    double last_t;
    int cx, cy, cz;
    int face;        // Face7 through which the current cube was entered
    uint32_t idx;    // linear index of the current cube in its volume (+ brick offset on the inner level)
};

AICB_DEV int signum_101(double x) { return (x == 0.0 || x != x) ? 0 : (x < 0.0 ? -1 : 1); }

// scale_to_integer_step (raycast.rs:797-819). fmod(s, 1) == s - trunc(s) exactly.
AICB_DEV double scale_to_integer_step(double s, double ds) {
    if (ds == 0.0 && !(s != s)) return __longlong_as_double(0x7ff0000000000000LL);
    if (ds < 0.0) {
        s = -s;
        ds = -ds;
    }
    double r = s - trunc(s);
    if (r < 0.0) r = r + 1.0;
    return (1.0 - r) / ds;
}

AICB_DEV bool in_i32_range(double x) { return (-2147483648.0 <= x) & (x < 2147483648.0); }

AICB_DEV double rclamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }  // NaN passes through

AICB_DEV float ps_mul(float a, float b) {
    float v = a * b;
    return (v != v) ? 0.0f : v;
}
AICB_DEV float ps_clamped(float v) { return (v > 0.0f) ? v : 0.0f; }
AICB_DEV float zo_clamped(float v) {
    if (v > 0.0f && v <= 1.0f) return v;
    if (v <= 0.0f) return 0.0f;
    return 1.0f;
}

// f32 transcendentals: evaluate in f64, round once (<= 1 ULP from glibc's powf/expf).
AICB_DEV float powf_exact(float x, float y) { return (float)pow((double)x, (double)y); }
AICB_DEV float expf_exact(float x) { return (float)exp((double)x); }

// Bounds of one level, in the form the stepping loop needs.
struct Level {
    int lo[3];
    int hi[3];     // exclusive
    int sy_sz;     // size_y * size_z
    int sz;        // size_z
    uint32_t base; // index offset (brick_off on the inner level)
};

// Raycaster::new(...).within(bounds, true) (raycast.rs:196-230, 513-545, 632-704) followed by the
// FirstLast::Beginning part of next() (raycast.rs:255-263): advance until the first in-bounds
// cube.  Returns false when the iterator produces nothing.
AICB_DEV bool caster_begin(Caster &c, const Ray &r, double ox, double oy, double oz, const Level &lv) {
    if (!(in_i32_range(ox) & in_i32_range(oy) & in_i32_range(oz))) return false;  // Cube::containing -> EMPTY
    {
        int fx = __double2int_rd(ox), fy = __double2int_rd(oy), fz = __double2int_rd(oz);
        const int lo = INT32_MIN + 1, hi = INT32_MAX - 1;
        if (fx < lo | fx >= hi | fy < lo | fy >= hi | fz < lo | fz >= hi) return false;  // MAXIMUM_BOUNDS filter
    }
    if (lv.hi[0] <= lv.lo[0] || lv.hi[1] <= lv.lo[1] || lv.hi[2] <= lv.lo[2]) return false;  // ORIGIN_EMPTY

    // fast_forward: (plane - origin) / direction per moving axis; the dot products with an axis
    // normal reduce exactly to this quotient.
    double max_t = 0.0;
    if (r.sx != 0) max_t = fmax(max_t, ((double)(r.sx < 0 ? lv.hi[0] : lv.lo[0]) - ox) / r.dx);
    if (r.sy != 0) max_t = fmax(max_t, ((double)(r.sy < 0 ? lv.hi[1] : lv.lo[1]) - oy) / r.dy);
    if (r.sz != 0) max_t = fmax(max_t, ((double)(r.sz < 0 ? lv.hi[2] : lv.lo[2]) - oz) / r.dz);

    double px = ox, py = oy, pz = oz, t0 = 0.0;
    if (max_t > 0.0) {
        double t_start = max_t - r.half_over_len;
        if (!isfinite(t_start)) t_start = max_t;
        px = ox + r.dx * t_start;  // Ray::advance (ray.rs:107-112)
        py = oy + r.dy * t_start;
        pz = oz + r.dz * t_start;
        if (!(in_i32_range(px) & in_i32_range(py) & in_i32_range(pz))) return false;
        t0 = t_start;
    }
    c.cx = __double2int_rd(px);
    c.cy = __double2int_rd(py);
    c.cz = __double2int_rd(pz);
    c.tmx = scale_to_integer_step(px, r.dx) + t0;
    c.tmy = scale_to_integer_step(py, r.dy) + t0;
    c.tmz = scale_to_integer_step(pz, r.dz) + t0;
    c.last_t = t0;
    c.face = AICB_FACE_WITHIN;

    // valid_for_stepping (raycast.rs:563-570)
    const bool any_nan = (c.tmx != c.tmx) | (c.tmy != c.tmy) | (c.tmz != c.tmz);
    const bool any_fin = isfinite(c.tmx) | isfinite(c.tmy) | isfinite(c.tmz);
    const bool valid = r.steppable & !any_nan & any_fin;

    for (;;) {
        // is_out_of_bounds_ahead (raycast.rs:711-728)
        bool xl = c.cx < lv.lo[0], xh = c.cx >= lv.hi[0];
        bool yl = c.cy < lv.lo[1], yh = c.cy >= lv.hi[1];
        bool zl = c.cz < lv.lo[2], zh = c.cz >= lv.hi[2];
        bool enter = (r.sx == 0 ? (xl | xh) : (r.sx < 0 ? xh : xl)) | (r.sy == 0 ? (yl | yh) : (r.sy < 0 ? yh : yl)) |
                     (r.sz == 0 ? (zl | zh) : (r.sz < 0 ? zh : zl));
        bool exit_ = (r.sx == 0 ? (xl | xh) : (r.sx < 0 ? xl : xh)) | (r.sy == 0 ? (yl | yh) : (r.sy < 0 ? yl : yh)) |
                     (r.sz == 0 ? (zl | zh) : (r.sz < 0 ? zl : zh));
        if (exit_) return false;
        if (!enter) break;
        if (!valid) return false;
        // State::step (raycast.rs:577-626)
        if (c.tmx < c.tmy) {
            if (c.tmx < c.tmz) { c.last_t = c.tmx; c.cx += r.sx; c.tmx += r.tdx; c.face = r.sx > 0 ? AICB_FACE_NX : AICB_FACE_PX; }
            else               { c.last_t = c.tmz; c.cz += r.sz; c.tmz += r.tdz; c.face = r.sz > 0 ? AICB_FACE_NZ : AICB_FACE_PZ; }
        } else {
            if (c.tmy < c.tmz) { c.last_t = c.tmy; c.cy += r.sy; c.tmy += r.tdy; c.face = r.sy > 0 ? AICB_FACE_NY : AICB_FACE_PY; }
            else               { c.last_t = c.tmz; c.cz += r.sz; c.tmz += r.tdz; c.face = r.sz > 0 ? AICB_FACE_NZ : AICB_FACE_PZ; }
        }
    }
    c.idx = lv.base + (uint32_t)((c.cx - lv.lo[0]) * lv.sy_sz + (c.cy - lv.lo[1]) * lv.sz + (c.cz - lv.lo[2]));
    // If stepping is impossible the iterator yields this one cube only when face == Within
    // (raycast.rs:245-249); with !valid no Beginning step can have happened, so it is.  The
    // caller learns about `valid` through Ray::steppable and the t_max test below.
    return true;
}

AICB_DEV bool caster_valid(const Caster &c, const Ray &r) {
    const bool any_nan = (c.tmx != c.tmx) | (c.tmy != c.tmy) | (c.tmz != c.tmz);
    const bool any_fin = isfinite(c.tmx) | isfinite(c.tmy) | isfinite(c.tmz);
    return r.steppable & !any_nan & any_fin;
}

// One State::step (raycast.rs:577-626) on the active caster, with incremental index update.
// Returns true if the new cube is outside the level (the "exit" step of raycast.rs:265-274).
AICB_DEV bool caster_step(Caster &c, const Ray &r, const Level &lv) {
    bool exited;
    if (c.tmx < c.tmy) {
        if (c.tmx < c.tmz) {
            c.last_t = c.tmx; c.cx += r.sx; c.tmx += r.tdx;
            c.face = r.sx > 0 ? AICB_FACE_NX : AICB_FACE_PX;
            c.idx += (uint32_t)(r.sx * lv.sy_sz);
            exited = r.sx > 0 ? (c.cx >= lv.hi[0]) : (c.cx < lv.lo[0]);
        } else {
            c.last_t = c.tmz; c.cz += r.sz; c.tmz += r.tdz;
            c.face = r.sz > 0 ? AICB_FACE_NZ : AICB_FACE_PZ;
            c.idx += (uint32_t)r.sz;
            exited = r.sz > 0 ? (c.cz >= lv.hi[2]) : (c.cz < lv.lo[2]);
        }
    } else {
        if (c.tmy < c.tmz) {
            c.last_t = c.tmy; c.cy += r.sy; c.tmy += r.tdy;
            c.face = r.sy > 0 ? AICB_FACE_NY : AICB_FACE_PY;
            c.idx += (uint32_t)(r.sy * lv.sz);
            exited = r.sy > 0 ? (c.cy >= lv.hi[1]) : (c.cy < lv.lo[1]);
        } else {
            c.last_t = c.tmz; c.cz += r.sz; c.tmz += r.tdz;
            c.face = r.sz > 0 ? AICB_FACE_NZ : AICB_FACE_PZ;
            c.idx += (uint32_t)r.sz;
            exited = r.sz > 0 ? (c.cz >= lv.hi[2]) : (c.cz < lv.lo[2]);
        }
    }
    return exited;
}

// RaycastStep::intersection_point (raycast.rs:409-439) for the caster's current (un-stepped)
// state, against the ray origin (ox,oy,oz) of that level.
AICB_DEV void intersection_point(const Caster &c, const Ray &r, double ox, double oy, double oz, double ip[3]) {
    if (c.face == AICB_FACE_WITHIN) {
        ip[0] = ox; ip[1] = oy; ip[2] = oz;
        return;
    }
    const int fa = (c.face - 1) % 3;
    const double tm[3] = {c.tmx, c.tmy, c.tmz};
    const double d[3] = {r.dx, r.dy, r.dz};
    const double o[3] = {ox, oy, oz};
    const int s[3] = {r.sx, r.sy, r.sz};
    const int cu[3] = {c.cx, c.cy, c.cz};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double p = (double)cu[a];
        if (a == fa) {
            if (s[a] < 0) p += 1.0;
        } else if (s[a] == 0) {
            p = o[a];
        } else {
            double off = (tm[a] - c.last_t) * d[a];
            p += (s[a] > 0) ? (1.0 - rclamp01(off)) : rclamp01(-off);
        }
        ip[a] = p;
    }
}

// ---- light ---------------------------------------------------------------------------------------
constexpr uint32_t TEXEL_ONE = 144u | (144u << 8) | (144u << 16) | (255u << 24);       // PackedLight::ONE
constexpr uint32_t TEXEL_NO_RAYS = (1u << 24);
constexpr uint32_t TEXEL_UNINIT = 0u;

// BlockSky::light_outside (sky.rs:113-147)
AICB_DEV uint32_t light_outside(const DeviceScene &s, int x, int y, int z) {
    const int c[3] = {x, y, z};
    int n_equal = 0, n_less = 0, which = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int beyond = s.lo[a] - 1;
        int hi = s.lo[a] + s.size[a];
        if (beyond == c[a]) { n_equal++; which = a; } else if (beyond < c[a]) n_less++;
        if (c[a] == hi) { n_equal++; which = 3 + a; } else if (c[a] < hi) n_less++;
    }
    if (n_less == 6) return TEXEL_UNINIT;
    if (n_equal == 1 && n_less == 5) return s.sky_faces[which];
    return TEXEL_NO_RAYS;
}

// SpaceRaytracer::get_packed_light (sr.rs:241-246)
AICB_DEV uint32_t get_packed_light(const DeviceScene &s, int x, int y, int z, uint32_t &texels) {
    uint32_t dx = (uint32_t)(x - s.lo[0]), dy = (uint32_t)(y - s.lo[1]), dz = (uint32_t)(z - s.lo[2]);
    if ((dx >= (uint32_t)s.size[0]) | (dy >= (uint32_t)s.size[1]) | (dz >= (uint32_t)s.size[2]))
        return light_outside(s, x, y, z);
    if (s.light == nullptr) return TEXEL_ONE;
    texels++;
    return __ldg(s.light + ((size_t)dx * s.size[1] + dy) * s.size[2] + dz);
}

AICB_DEV void texel_value_ao(const DeviceScene &s, uint32_t t, float out[4]) {  // data.rs:145-158
    out[0] = __ldg(s.lut + (t & 255));
    out[1] = __ldg(s.lut + ((t >> 8) & 255));
    out[2] = __ldg(s.lut + ((t >> 16) & 255));
    uint32_t st = t >> 24;
    out[3] = (st == 255) ? 1.0f : (st == 128 ? 0.25f : 0.0f);
}

AICB_DEV double rem_euclid1(double x) {
    double r = x - trunc(x);
    return r < 0.0 ? r + 1.0 : r;
}

// get_interpolated_light (sr.rs:248-359)
AICB_DEV void interpolated_light(const DeviceScene &s, uint32_t mode, int cube_x, int cube_y, int cube_z, int face,
                                 const double sp[3], float out[3], uint32_t &texels) {
    const double eps = 0.5 / 256.0;
    // Face::rotation_from_nz (face.rs:395-405): axis + sign of the images of +X and +Y
    int a1, s1, a2, s2, an, sn;
    switch (face) {
        case AICB_FACE_NX: a1 = 1; s1 = 1;  a2 = 2; s2 = 1;  an = 0; sn = -1; break;  // RYZX
        case AICB_FACE_NY: a1 = 2; s1 = 1;  a2 = 0; s2 = 1;  an = 1; sn = -1; break;  // RZXY
        case AICB_FACE_NZ: a1 = 0; s1 = 1;  a2 = 1; s2 = 1;  an = 2; sn = -1; break;  // RXYZ
        case AICB_FACE_PX: a1 = 1; s1 = -1; a2 = 2; s2 = 1;  an = 0; sn = 1;  break;  // RyZx
        case AICB_FACE_PY: a1 = 2; s1 = 1;  a2 = 0; s2 = -1; an = 1; sn = 1;  break;  // RZxy
        case AICB_FACE_PZ: a1 = 0; s1 = 1;  a2 = 1; s2 = -1; an = 2; sn = 1;  break;  // RXyz
        default:           a1 = 0; s1 = 1;  a2 = 1; s2 = 1;  an = 2; sn = 0;  break;  // Within: IDENTITY, normal 0
    }
    double mix_1 = rem_euclid1((s1 > 0 ? sp[a1] : -sp[a1]) - 0.5);
    double mix_2 = rem_euclid1((s2 > 0 ? sp[a2] : -sp[a2]) - 0.5);
    if (mix_1 > 0.5) { mix_1 = 1.0 - mix_1; s1 = -s1; }
    if (mix_2 > 0.5) { mix_2 = 1.0 - mix_2; s2 = -s2; }
    if (mode == AICB_LIGHT_COARSE) {          // surface.rs:510-514
        double f1 = floor(mix_1 * 4.0), f2 = floor(mix_2 * 4.0);
        f1 = f1 < 0.0 ? 0.0 : (f1 > 3.0 ? 3.0 : f1);
        f2 = f2 < 0.0 ? 0.0 : (f2 > 3.0 ? 3.0 : f2);
        mix_1 = (f1 + 0.5) / 4.0;
        mix_2 = (f2 + 0.5) / 4.0;
    } else if (mode == AICB_LIGHT_SMOOTHSTEP) {  // surface.rs:517-520
        double c1 = rclamp01(mix_1), c2 = rclamp01(mix_2);
        mix_1 = 3.0 * (c1 * c1) - 2.0 * ((c1 * c1) * c1);
        mix_2 = 3.0 * (c2 * c2) - 2.0 * ((c2 * c2) * c2);
    }
    const float m1 = (float)mix_1, m2 = (float)mix_2;

    const int cube[3] = {cube_x, cube_y, cube_z};
    double fdot_sp = sn == 0 ? 0.0 : (sn > 0 ? sp[an] : -sp[an]);
    double ctr = (double)cube[an] + 0.5;
    double fdot_c = sn == 0 ? 0.0 : (sn > 0 ? ctr : -ctr);
    const double height_in_cube = fdot_sp - fdot_c + 0.5;

    const double lo1 = (double)s1 * -0.5, hi1 = (double)s1 * 0.5;
    const double lo2 = (double)s2 * -0.5, hi2 = (double)s2 * 0.5;

    float front[4], result[4];
#pragma unroll 1
    for (int layer = 0; layer < 2; layer++) {
        const double along = (layer == 0) ? (1.0 - eps) : eps;
        double p[3] = {sp[0], sp[1], sp[2]};
        if (sn != 0) p[an] = sp[an] + (double)sn * along;
        const double b1 = p[a1], b2 = p[a2];
        uint32_t tex[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // k: 0 near12, 1 near1far2, 2 near2far1, 3 far12
            p[a1] = b1 + ((k & 2) ? hi1 : lo1);
            p[a2] = b2 + ((k & 1) ? hi2 : lo2);
            if (in_i32_range(p[0]) & in_i32_range(p[1]) & in_i32_range(p[2]))
                tex[k] = get_packed_light(s, __double2int_rd(p[0]), __double2int_rd(p[1]), __double2int_rd(p[2]), texels);
            else
                tex[k] = s.sky_mean;
        }
        if ((tex[1] >> 24) != 255 && (tex[2] >> 24) != 255) tex[3] = tex[0];  // sr.rs:317-321
        float v0[4], v1[4], v2[4], v3[4], cur[4];
        texel_value_ao(s, tex[0], v0);
        texel_value_ao(s, tex[1], v1);
        texel_value_ao(s, tex[2], v2);
        texel_value_ao(s, tex[3], v3);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float ab = v0[i] + (v1[i] - v0[i]) * m2;
            float cd = v2[i] + (v3[i] - v2[i]) * m2;
            cur[i] = ab + (cd - ab) * m1;
        }
        if (layer == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) { front[i] = cur[i]; result[i] = cur[i]; }
            if (height_in_cube > (1.0 - eps)) break;
        } else {
            const float h = (float)height_in_cube;
#pragma unroll
            for (int i = 0; i < 4; i++) result[i] = cur[i] + (front[i] - cur[i]) * h;
        }
    }
    const float w = fmaxf(result[3], 0.1f);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float v = result[i] / w;
        out[i] = (v == 0.0f) ? 0.0f : v;
    }
}

// ---- per-ray tracing state -------------------------------------------------------------------------
template <bool AUX>
struct AuxState {};
template <>
struct AuxState<true> {
    double depth;
    int hit_cube[3];
    int hit_voxel[3];
    int hit_res;
    int hit_face;
    bool have_hit;
    uint32_t n_outer, n_inner, n_hits, n_texels, n_blocks;
};

// A surface remembered between discovery and shading (Volumetric mode pairs it with the next
// event's t, surface.rs:460-490).  Illumination depends only on geometry, so it is evaluated at
// discovery and carried as three floats instead of carrying the intersection point.
template <int LC, bool AUX>
struct PendingSurface {
    uint32_t pal;      // global palette entry index
    double t;
    float illum[3];
    int cube[3];
    uint32_t packed;   // voxel x | y<<8 | z<<16 | face<<24 ; resolution in hit_res
    int res;
};

template <bool VOLUMETRIC, int LC, bool AUX>
struct Tracer {
    const TraceParams &P;
    float lr, lg, lb, T;          // ColorBuf (raytracer_components.rs:20-39)
    uint32_t steps;               // primary_cubes_traced (sr.rs:612)
    double t_to_abs;              // sr.rs:146
    float t_to_view;              // sr.rs:149-151
    bool have_fog;
    float fog_r, fog_g, fog_b, fog_blend;
    AuxState<AUX> aux;
    bool have_last;
    PendingSurface<LC, AUX> last;

    __device__ explicit Tracer(const TraceParams &p) : P(p) {}

    // count_step_should_stop (sr.rs:625-656); the EnterSpace / Incomplete hits are no-ops for ColorBuf
    AICB_DEV bool count_stop() {
        steps += 1;
        if (steps > 1000) return true;
        return T < (1.0f / 256.0f);
    }

    // distance_fog (sr.rs:745-768)
    AICB_DEV float fog_amount(double t) const {
        float rel = (float)t * t_to_view;
        rel = rel < 0.0f ? 0.0f : (rel > 1.0f ? 1.0f : rel);
        float fog_exponential = 1.0f - expf_exact(-1.6f * rel);
        float fudged = fog_exponential / 0.79810348f;
        float p4 = (rel * rel) * (rel * rel);
        return zo_clamped(fudged * (1.0f - fog_blend) + p4 * fog_blend);
    }

    // Surface::to_light + trace_through_surface (surface.rs:73-106, sr.rs:697-717)
    AICB_DEV void shade(float cr, float cg, float cb, float ca, float er, float eg, float eb,
                        const PendingSurface<LC, AUX> &sf) {
        if (P.transparency == AICB_TRANSPARENCY_THRESHOLD) {  // limit_alpha (graphics_options.rs:496-507)
            if (ca > P.threshold) { ca = 1.0f; } else { cr = cg = cb = ca = 0.0f; }
        }
        if (ca == 0.0f && er == 0.0f && eg == 0.0f && eb == 0.0f) return;
        float ir = 1.0f, ig = 1.0f, ib = 1.0f;
        if (LC != LC_NONE) { ir = sf.illum[0]; ig = sf.illum[1]; ib = sf.illum[2]; }
        float orr = ps_mul(ps_mul(cr, ir), ca) + er;   // reflect + emission (color.rs:708-710)
        float og = ps_mul(ps_mul(cg, ig), ca) + eg;
        float ob = ps_mul(ps_mul(cb, ib), ca) + eb;
        float tr = 1.0f - ca;
        if (have_fog) {
            float fa = fog_amount(sf.t);
            float comp = 1.0f - fa;
            orr = ps_mul(orr, comp) + ps_mul(fog_r, fa);
            og = ps_mul(og, comp) + ps_mul(fog_g, fa);
            ob = ps_mul(ob, comp) + ps_mul(fog_b, fa);
            tr = tr * comp;
        }
        // add_color_internal (raytracer_components.rs:87-92)
        lr = lr + orr * T;
        lg = lg + og * T;
        lb = lb + ob * T;
        T = T * tr;
        if (AUX) {
            aux_hit(sf);
        }
    }

    AICB_DEV void aux_hit(const PendingSurface<LC, AUX> &sf) {
        if constexpr (AUX) {
            aux.depth = fmin(aux.depth, sf.t);
            aux.n_hits++;
            if (!aux.have_hit) {
                aux.have_hit = true;
                aux.hit_cube[0] = sf.cube[0]; aux.hit_cube[1] = sf.cube[1]; aux.hit_cube[2] = sf.cube[2];
                aux.hit_voxel[0] = sf.packed & 255; aux.hit_voxel[1] = (sf.packed >> 8) & 255;
                aux.hit_voxel[2] = (sf.packed >> 16) & 255;
                aux.hit_face = sf.packed >> 24;
                aux.hit_res = sf.res;
            }
        }
    }

    AICB_DEV void shade_pending(const PendingSurface<LC, AUX> &sf) {
        float4 c = __ldg(P.scene.palette + 2 * (size_t)sf.pal);
        float4 e = __ldg(P.scene.palette + 2 * (size_t)sf.pal + 1);
        shade(c.x, c.y, c.z, c.w, e.x, e.y, e.z, sf);
    }

    // trace_through_span (sr.rs:720-740) with apply_transmittance (raytracer_components.rs:215-258)
    AICB_DEV void shade_span(const PendingSurface<LC, AUX> &sf, double exit_t) {
        float4 c = __ldg(P.scene.palette + 2 * (size_t)sf.pal);
        float4 e = __ldg(P.scene.palette + 2 * (size_t)sf.pal + 1);
        float thickness = fmaxf((float)((exit_t - sf.t) * t_to_abs), 0.0f);
        float alpha, coeff;
        float cr = c.x, cg = c.y, cb = c.z;
        if (thickness == 0.0f) {
            if (c.w == 1.0f) { alpha = c.w; coeff = 1.0f; }
            else { cr = cg = cb = 0.0f; alpha = 0.0f; coeff = 0.0f; }
        } else {
            float unit_t = 1.0f - c.w;
            float depth_t = powf_exact(unit_t, thickness);
            alpha = zo_clamped(1.0f - depth_t);
            float k = (unit_t == 1.0f) ? thickness : (depth_t - 1.0f) / (unit_t - 1.0f);
            coeff = fmaxf(k, 0.0f);
        }
        float k = ps_clamped(coeff);
        shade(cr, cg, cb, alpha, ps_mul(e.x, k), ps_mul(e.y, k), ps_mul(e.z, k), sf);
    }

    // compute_illumination (surface.rs:113-206) at discovery time
    AICB_DEV void illuminate(PendingSurface<LC, AUX> &sf, const double ip[3]) {
        if constexpr (LC == LC_FLAT) {
            int x = sf.cube[0], y = sf.cube[1], z = sf.cube[2];
            int face = sf.packed >> 24;
            if (face != AICB_FACE_WITHIN) {
                int d = face >= AICB_FACE_PX ? 1 : -1;
                int ax = (face - 1) % 3;
                if (ax == 0) x += d; else if (ax == 1) y += d; else z += d;
            }
            uint32_t tx = 0;
            uint32_t t = get_packed_light(P.scene, x, y, z, tx);
            if constexpr (AUX) aux.n_texels += tx;
            sf.illum[0] = __ldg(P.scene.lut + (t & 255));
            sf.illum[1] = __ldg(P.scene.lut + ((t >> 8) & 255));
            sf.illum[2] = __ldg(P.scene.lut + ((t >> 16) & 255));
        } else if constexpr (LC == LC_INTERP) {
            uint32_t tx = 0;
            interpolated_light(P.scene, P.lighting, sf.cube[0], sf.cube[1], sf.cube[2], sf.packed >> 24, ip, sf.illum, tx);
            if constexpr (AUX) aux.n_texels += tx;
        }
    }

    // One TraceStep through the Surface-mode loop (sr.rs:206-225) or DepthIter + the Volumetric
    // loop (surface.rs:460-490, sr.rs:185-203).  kind: 0 EnterSurface, 1 Invisible, 2 EnterBlock.
    // Returns true when tracing must stop.
    AICB_DEV bool process(int kind, double t, const PendingSurface<LC, AUX> &sf) {
        if constexpr (!VOLUMETRIC) {
            if (count_stop()) return true;
            if (kind == 0) shade_pending(sf);
            return false;
        } else {
            const bool emit_span = have_last;
            PendingSurface<LC, AUX> span = last;
            double exit_t = t;
            if (kind == 0) {
                exit_t = sf.t;
                last = sf;
                have_last = true;
            } else {
                have_last = false;
            }
            if (count_stop()) return true;
            if (emit_span) shade_span(span, exit_t);
            if (kind == 2) {
                if (count_stop()) return true;  // the buffered DepthStep::EnterBlock
            }
            return false;
        }
    }

    // trace_ray_impl (sr.rs:135-238)
    __device__ void trace(double ox, double oy, double oz, double dx, double dy, double dz) {
        const DeviceScene &S = P.scene;
        lr = lg = lb = 0.0f;
        T = 1.0f;
        steps = 0;
        have_last = false;
        if constexpr (AUX) {
            aux.depth = __longlong_as_double(0x7ff0000000000000LL);
            aux.have_hit = false;
            aux.n_outer = aux.n_inner = aux.n_hits = aux.n_texels = aux.n_blocks = 0;
        }

        // Sky::sample (sky.rs:32-41)
        float sky_r = 0.0f, sky_g = 0.0f, sky_b = 0.0f;
        if (P.include_sky) {
            int k = 0;
            if (S.sky_kind) k = ((dx >= 0.0) << 2) + ((dy >= 0.0) << 1) + (dz >= 0.0);
            sky_r = S.sky_colors[k][0]; sky_g = S.sky_colors[k][1]; sky_b = S.sky_colors[k][2];
        }
        t_to_abs = sqrt(dx * dx + dy * dy + dz * dz);
        t_to_view = (float)(t_to_abs / P.view_distance);
        have_fog = (P.fog != AICB_FOG_NONE) && P.include_sky;
        fog_r = sky_r; fog_g = sky_g; fog_b = sky_b;
        fog_blend = (P.fog == AICB_FOG_ABRUPT) ? 1.0f : (P.fog == AICB_FOG_COMPROMISE ? 0.5f : 0.0f);

        // Parameters::new (raycast.rs:749-771)
        Ray r;
        r.ox = ox; r.oy = oy; r.oz = oz;
        if (!((fabs(dx) < 1e100) & (fabs(dy) < 1e100) & (fabs(dz) < 1e100))) { dx = dy = dz = 0.0; }
        r.dx = dx; r.dy = dy; r.dz = dz;
        r.sx = signum_101(dx); r.sy = signum_101(dy); r.sz = signum_101(dz);
        r.tdx = 1.0 / fabs(dx); r.tdy = 1.0 / fabs(dy); r.tdz = 1.0 / fabs(dz);
        r.half_over_len = 0.5 / sqrt(dx * dx + dy * dy + dz * dz);
        r.steppable = (r.sx | r.sy | r.sz) != 0;

        Level outer;
        outer.lo[0] = S.lo[0]; outer.lo[1] = S.lo[1]; outer.lo[2] = S.lo[2];
        outer.hi[0] = S.lo[0] + S.size[0]; outer.hi[1] = S.lo[1] + S.size[1]; outer.hi[2] = S.lo[2] + S.size[2];
        outer.sy_sz = S.size[1] * S.size[2];
        outer.sz = S.size[2];
        outer.base = 0;

        Caster c;
        bool running = caster_begin(c, r, ox, oy, oz, outer);
        bool valid = running && caster_valid(c, r);

        // level state
        bool inner = false;
        Level lv = outer;
        Caster saved;            // outer caster while inside a block
        bool saved_valid = false;
        double sub_ox = 0, sub_oy = 0, sub_oz = 0, antiscale = 1.0;
        uint32_t pal_off = 0;
        int res = 1;
        bool need_advance = false;

        while (running) {
            if (need_advance) {
                bool exited;
                if (!valid) {
                    // cannot step: the iterator ends without an exit step (raycast.rs:245-249)
                    exited = true;
                    if (inner) { inner = false; c = saved; lv = outer; valid = saved_valid; need_advance = true; continue; }
                    break;
                }
                exited = caster_step(c, r, lv);
                if (exited) {
                    // exit step: Invisible at this t (surface.rs:296-301, 388-393)
                    PendingSurface<LC, AUX> none;
                    if (process(1, inner ? c.last_t * antiscale : c.last_t, none)) break;
                    if (inner) { inner = false; c = saved; lv = outer; valid = saved_valid; need_advance = true; continue; }
                    break;
                }
            }
            need_advance = true;

            // ---- look up the current cube / voxel --------------------------------------------
            PendingSurface<LC, AUX> sf;
            int kind;  // TraceStep kind
            double t;
            if (!inner) {
                if constexpr (AUX) aux.n_outer++;
                uint32_t cell = S.wide_cells ? __ldg((const uint32_t *)S.cells + c.idx)
                                             : (uint32_t)__ldg((const uint16_t *)S.cells + c.idx);
                uint32_t ck = S.wide_cells ? (cell >> 16) : (cell >> 14);
                uint32_t id = S.wide_cells ? (cell & 0xffffu) : (cell & 0x3fffu);
                t = c.last_t;
                if (ck == KIND_INVISIBLE) {
                    kind = 1;
                } else {
                    const uint4 *bp = reinterpret_cast<const uint4 *>(S.blocks + id);
                    uint4 b0 = __ldg(bp);
                    uint4 b1 = __ldg(bp + 1);
                    if (ck == KIND_SINGLE) {
                        kind = 0;
                        sf.pal = b1.y;
                        sf.t = t;
                        sf.cube[0] = c.cx; sf.cube[1] = c.cy; sf.cube[2] = c.cz;
                        sf.packed = (uint32_t)c.face << 24;
                        sf.res = 1;
                        if constexpr (LC == LC_INTERP) {
                            double ip[3];
                            intersection_point(c, r, r.ox, r.oy, r.oz, ip);
                            illuminate(sf, ip);
                        } else if constexpr (LC == LC_FLAT) {
                            double ip[3] = {0, 0, 0};
                            illuminate(sf, ip);
                        }
                    } else {
                        // recursive_raycast (raycast.rs:458-476) + TraceStep::EnterBlock (surface.rs:334-352)
                        if constexpr (AUX) aux.n_blocks++;
                        PendingSurface<LC, AUX> none;
                        if (process(2, t, none)) break;
                        res = (int)(b0.x >> 8);
                        Level in;
                        in.lo[0] = (int16_t)(b0.y & 0xffff); in.lo[1] = (int16_t)(b0.y >> 16); in.lo[2] = (int16_t)(b0.z & 0xffff);
                        int vsx = (int)(b0.z >> 16), vsy = (int)(b0.w & 0xffff), vsz = (int)(b0.w >> 16);
                        in.hi[0] = in.lo[0] + vsx; in.hi[1] = in.lo[1] + vsy; in.hi[2] = in.lo[2] + vsz;
                        in.sy_sz = vsy * vsz;
                        in.sz = vsz;
                        in.base = b1.x;
                        double fres = (double)res;
                        double so_x = (r.ox - (double)c.cx) * fres;
                        double so_y = (r.oy - (double)c.cy) * fres;
                        double so_z = (r.oz - (double)c.cz) * fres;
                        Caster ic;
                        if (caster_begin(ic, r, so_x, so_y, so_z, in)) {
                            saved = c;
                            saved_valid = valid;
                            c = ic;
                            lv = in;
                            valid = caster_valid(c, r);
                            inner = true;
                            sub_ox = so_x; sub_oy = so_y; sub_oz = so_z;
                            antiscale = 1.0 / fres;
                            pal_off = b1.y;
                            need_advance = false;
                        }
                        continue;
                    }
                }
            } else {
                if constexpr (AUX) aux.n_inner++;
                uint32_t v = __ldg(S.bricks + c.idx);
                t = c.last_t * antiscale;  // surface.rs:385-386
                if (v & 0x8000u) {
                    kind = 1;
                } else {
                    kind = 0;
                    sf.pal = pal_off + v;
                    sf.t = t;
                    sf.cube[0] = saved.cx; sf.cube[1] = saved.cy; sf.cube[2] = saved.cz;
                    sf.packed = (uint32_t)c.cx | ((uint32_t)c.cy << 8) | ((uint32_t)c.cz << 16) | ((uint32_t)c.face << 24);
                    sf.res = res;
                    if constexpr (LC == LC_INTERP) {
                        double ip[3];
                        intersection_point(c, r, sub_ox, sub_oy, sub_oz, ip);
                        ip[0] = ip[0] * antiscale + (double)saved.cx;  // surface.rs:406-407
                        ip[1] = ip[1] * antiscale + (double)saved.cy;
                        ip[2] = ip[2] * antiscale + (double)saved.cz;
                        illuminate(sf, ip);
                    } else if constexpr (LC == LC_FLAT) {
                        double ip[3] = {0, 0, 0};
                        illuminate(sf, ip);
                    }
                }
            }
            if (process(kind, t, sf)) break;
        }

        // finish (sr.rs:658-693): the sky is an opaque hit at t = inf
        if (P.include_sky) {
            lr = lr + (sky_r * 1.0f) * T;
            lg = lg + (sky_g * 1.0f) * T;
            lb = lb + (sky_b * 1.0f) * T;
            T = T * (1.0f - 1.0f);
        }
        if (P.debug_pixel_cost) {
            // ColorBuf::add for Exception::DebugOverrideRg (accum.rs:228-234)
            float k = ps_clamped((float)steps);
            float red = ps_clamped(ps_mul(0.02f, k) * 1.0f);
            float green = ps_clamped(ps_mul(0.002f, k) * 1.0f);
            float rgba[4];
            colorbuf_to_rgba(lr, lg, lb, T, rgba);
            float lum = rgba[1] * 0.7152f + (rgba[0] * 0.2126f + rgba[2] * 0.0722f);
            lr = red; lg = green; lb = ps_clamped(lum * 0.2f);
            T = 0.0f;
        }
    }

    // Rgba::from(ColorBuf) (raytracer_components.rs:122-146)
    static AICB_DEV void colorbuf_to_rgba(float l0, float l1, float l2, float tr, float out[4]) {
        if (tr >= 1.0f) { out[0] = out[1] = out[2] = out[3] = 0.0f; return; }
        float alpha = 1.0f - tr;
        float c[3] = {l0 / alpha, l1 / alpha, l2 / alpha};
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (c[i] > 0.0f) {} else if (c[i] == 0.0f) c[i] = 0.0f; else ok = false;
        }
        if (!ok) { c[0] = 1.0f; c[1] = 0.0f; c[2] = 0.0f; }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
        out[3] = (alpha > 0.0f && alpha <= 1.0f) ? alpha : (alpha == 0.0f ? 0.0f : 1.0f);
    }
};

// component_to_srgb8 (color.rs:1038-1054): `as u8` saturates, round() is half-away-from-zero
AICB_DEV unsigned char sat_u8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (unsigned char)v;
}
AICB_DEV unsigned char component_to_srgb8(float c) {
    float s = (c <= 0.0031308f) ? c * (323.0f / 25.0f) : (211.0f * powf_exact(c, 5.0f / 12.0f) - 11.0f) / 200.0f;
    return sat_u8(roundf(s * 255.0f));
}

// Camera::post_process_color + to_srgb8 (camera_struct.rs:376-382, graphics_options.rs:352-368, color.rs:669-676)
AICB_DEV uchar4 encode_srgb8(const TraceParams &P, float l0, float l1, float l2, float tr) {
    float rgba[4];
    Tracer<false, LC_NONE, false>::colorbuf_to_rgba(l0, l1, l2, tr, rgba);
    float c[3] = {ps_mul(rgba[0], P.exposure), ps_mul(rgba[1], P.exposure), ps_mul(rgba[2], P.exposure)};
    if (isfinite(P.maximum_intensity)) {
        if (P.tone_mapping == AICB_TONE_CLAMP) {
#pragma unroll
            for (int i = 0; i < 3; i++) c[i] = c[i] > P.maximum_intensity ? P.maximum_intensity : c[i];
        } else {
            float lum = c[1] * 0.7152f + (c[0] * 0.2126f + c[2] * 0.0722f);
            float s = ps_clamped(1.0f / (1.0f + lum / P.maximum_intensity));
#pragma unroll
            for (int i = 0; i < 3; i++) c[i] = ps_mul(c[i], s);
        }
    }
    return make_uchar4(component_to_srgb8(c[0]), component_to_srgb8(c[1]), component_to_srgb8(c[2]),
                       sat_u8(roundf(rgba[3] * 255.0f)));
}

// Camera::project_ndc_into_world (camera_struct.rs:238-257); euclid transform_point3d
AICB_DEV void project_ndc(const TraceParams &P, double x, double y, double z, double out[3]) {
    const double *m = P.m;
    double hx = x * m[0] + y * m[4] + z * m[8] + m[12];
    double hy = x * m[1] + y * m[5] + z * m[9] + m[13];
    double hz = x * m[2] + y * m[6] + z * m[10] + m[14];
    double hw = x * m[3] + y * m[7] + z * m[11] + m[15];
    if (hw > 0.0) {
        out[0] = hx / hw; out[1] = hy / hw; out[2] = hz / hw;
    } else {
        out[0] = out[1] = out[2] = __longlong_as_double(0x7ff8000000000000LL);
    }
}

// viewport.rs:104-113 + renderer.rs:424-451,489-491
AICB_DEV void pixel_ray(const TraceParams &P, uint32_t xch, uint32_t ych, int sample, double o[3], double d[3]) {
    const double W = (double)P.fb_width, H = (double)P.fb_height;
    const double x0 = (double)xch / W * 2.0 - 1.0;
    const double x1 = (double)(xch + 1) / W * 2.0 - 1.0;
    const double y0 = -((double)ych / H * 2.0 - 1.0);
    const double y1 = -((double)(ych + 1) / H * 2.0 - 1.0);
    double px, py;
    if (sample < 0) {
        px = (x0 + x1) / 2.0;
        py = (y0 + y1) / 2.0;
    } else {
        const double u = (sample == 0) ? 1. / 8. : (sample == 1) ? 3. / 8. : (sample == 2) ? 5. / 8. : 7. / 8.;
        const double v = (sample == 0) ? 5. / 8. : (sample == 1) ? 1. / 8. : (sample == 2) ? 7. / 8. : 3. / 8.;
        px = x0 + (x1 - x0) * u;
        py = y0 + (y1 - y0) * v;
    }
    double nearp[3], farp[3];
    project_ndc(P, px, py, 0.0, nearp);
    project_ndc(P, px, py, 1.0, farp);
    o[0] = nearp[0]; o[1] = nearp[1]; o[2] = nearp[2];
    d[0] = farp[0] - nearp[0]; d[1] = farp[1] - nearp[1]; d[2] = farp[2] - nearp[2];
}

constexpr int TILE_W = 8, TILE_H = 4;
constexpr int WARPS_PER_BLOCK = 4;

// The frame kernel: persistent warps, one 8x4 pixel tile per warp at a time
// (replaces the Rayon dispatch trace_scene_to_image_impl, renderer.rs:516-556).
template <bool VOLUMETRIC, int LC, bool AUX>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
trace_kernel(const __grid_constant__ TraceParams P) {
    const int lane = threadIdx.x & 31;
    unsigned long long cubes_traced = 0;
    unsigned long long n_outer = 0, n_inner = 0, n_hits = 0, n_texels = 0, n_blocks = 0;
    const uint32_t n_tiles = P.tiles_x * P.tiles_y;
    const bool explicit_rays = P.rays != nullptr;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;

        uint32_t lx, ly_local;  // pixel within the (local) image
        bool active;
        size_t out_index;
        uint32_t gy = 0;
        if (explicit_rays) {
            size_t i = (size_t)tile * 32 + lane;
            active = i < P.n_rays;
            out_index = i;
            lx = ly_local = 0;
        } else {
            const uint32_t tx = tile % P.tiles_x, ty = tile / P.tiles_x;
            lx = tx * TILE_W + (lane & (TILE_W - 1));
            ly_local = ty * TILE_H + (lane / TILE_W);
            active = lx < P.fb_width && ly_local < P.local_rows;
            out_index = (size_t)ly_local * P.fb_width + lx;
            // local row -> framebuffer row (row-strip sharding)
            if (P.shard_count > 1) {
                uint32_t strip_local = ly_local / P.strip_rows;
                gy = (strip_local * P.shard_count + P.shard_index) * P.strip_rows + ly_local % P.strip_rows;
            } else {
                gy = ly_local;
            }
        }
        if (!active) continue;

        Tracer<VOLUMETRIC, LC, AUX> tr(P);
        float l0, l1, l2, tT;
        uint32_t steps_total = 0;
        double depth = 0;
        if (explicit_rays) {
            const double *rp = P.rays + 6 * out_index;
            tr.trace(rp[0], rp[1], rp[2], rp[3], rp[4], rp[5]);
            l0 = tr.lr; l1 = tr.lg; l2 = tr.lb; tT = tr.T;
            steps_total = tr.steps;
        } else if (!P.antialias) {
            double o[3], d[3];
            pixel_ray(P, lx, gy, -1, o, d);
            tr.trace(o[0], o[1], o[2], d[0], d[1], d[2]);
            l0 = tr.lr; l1 = tr.lg; l2 = tr.lb; tT = tr.T;
            steps_total = tr.steps;
        } else {
            // 4 fixed sub-samples, ColorBuf::mean (renderer.rs:426-444, raytracer_components.rs:97-102)
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, sT = 0.0f;
            AuxState<AUX> first_aux;
            bool got = false;
            double dmin = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll 1
            for (int k = 0; k < 4; k++) {
                double o[3], d[3];
                pixel_ray(P, lx, gy, k, o, d);
                tr.trace(o[0], o[1], o[2], d[0], d[1], d[2]);
                s0 = s0 + tr.lr; s1 = s1 + tr.lg; s2 = s2 + tr.lb; sT = sT + tr.T;
                steps_total += tr.steps;
                if constexpr (AUX) {
                    dmin = fmin(dmin, tr.aux.depth);
                    n_outer += tr.aux.n_outer; n_inner += tr.aux.n_inner; n_hits += tr.aux.n_hits;
                    n_texels += tr.aux.n_texels; n_blocks += tr.aux.n_blocks;
                    if (!got && (tr.aux.have_hit || k == 0)) { first_aux = tr.aux; got = tr.aux.have_hit; }
                }
            }
            l0 = s0 / 4.0f; l1 = s1 / 4.0f; l2 = s2 / 4.0f; tT = sT / 4.0f;
            if constexpr (AUX) {
                tr.aux = first_aux;
                tr.aux.depth = dmin;
                tr.aux.n_outer = tr.aux.n_inner = tr.aux.n_hits = tr.aux.n_texels = tr.aux.n_blocks = 0;
            }
        }
        cubes_traced += steps_total;
        (void)depth;

        if (P.out_srgb8) P.out_srgb8[out_index] = encode_srgb8(P, l0, l1, l2, tT);
        if (P.out_colorbuf) P.out_colorbuf[out_index] = make_float4(l0, l1, l2, tT);
        if constexpr (AUX) {
            n_outer += tr.aux.n_outer; n_inner += tr.aux.n_inner; n_hits += tr.aux.n_hits;
            n_texels += tr.aux.n_texels; n_blocks += tr.aux.n_blocks;
            if (P.out_depth) P.out_depth[out_index] = tr.aux.depth;
            if (P.out_steps) P.out_steps[out_index] = steps_total;
            if (P.out_hit) {
                aicb_hit h;
                if (tr.aux.have_hit) {
                    h.cube[0] = tr.aux.hit_cube[0]; h.cube[1] = tr.aux.hit_cube[1]; h.cube[2] = tr.aux.hit_cube[2];
                    h.voxel[0] = tr.aux.hit_voxel[0]; h.voxel[1] = tr.aux.hit_voxel[1]; h.voxel[2] = tr.aux.hit_voxel[2];
                    h.resolution = tr.aux.hit_res;
                    h.face = tr.aux.hit_face;
                } else {
                    h.cube[0] = h.cube[1] = h.cube[2] = -1;
                    h.voxel[0] = h.voxel[1] = h.voxel[2] = -1;
                    h.resolution = -1;
                    h.face = -1;
                }
                P.out_hit[out_index] = h;
            }
        }
    }

    // RaytraceInfo sum (renderer.rs:555): warp-reduce then one atomic per warp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cubes_traced += __shfl_down_sync(0xffffffffu, cubes_traced, off);
    if (lane == 0 && cubes_traced) atomicAdd(P.counters + 0, cubes_traced);
    if constexpr (AUX) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            n_outer += __shfl_down_sync(0xffffffffu, n_outer, off);
            n_inner += __shfl_down_sync(0xffffffffu, n_inner, off);
            n_hits += __shfl_down_sync(0xffffffffu, n_hits, off);
            n_texels += __shfl_down_sync(0xffffffffu, n_texels, off);
            n_blocks += __shfl_down_sync(0xffffffffu, n_blocks, off);
        }
        if (lane == 0) {
            atomicAdd(P.counters + 1, n_outer);
            atomicAdd(P.counters + 2, n_inner);
            atomicAdd(P.counters + 3, n_hits);
            atomicAdd(P.counters + 4, n_texels);
            atomicAdd(P.counters + 5, n_blocks);
        }
    }
}

#endif  // __CUDACC__

}  // namespace aicb
