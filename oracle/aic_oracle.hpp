// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17) of the all-is-cubes per-pixel voxel raytracer hot path, written
// from the reference's Rust sources (which cannot be compiled here: no rustc/cargo).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this.  The product (libaicb200.so) never links, includes or calls anything here.
//
// Parity pinning: every function cites the reference file:line it follows, and
// tests/test_oracle_*.py check it against the reference's own known-answer tests
// (raycast/tests.rs, raytracer/surface.rs, accum.rs, text.rs, raytracer_components.rs,
// camera/tests.rs, space/light/data.rs).
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math (Rust never contracts to FMA).
#pragma once
#include <cstddef>
#include <cstdint>

#include "../include/aicb200.h"

namespace orc {

constexpr int32_t I32_MIN = INT32_MIN;
constexpr int32_t I32_MAX = INT32_MAX;

// GridAab with exclusive upper bound. (all-is-cubes-base/src/math/grid_aab.rs)
struct Aab {
    int32_t lo[3];
    int32_t hi[3];
    bool empty() const { return hi[0] <= lo[0] || hi[1] <= lo[1] || hi[2] <= lo[2]; }
};

// raycast.rs:485-499
inline Aab maximum_bounds() {
    return Aab{{I32_MIN + 1, I32_MIN + 1, I32_MIN + 1}, {I32_MAX - 1, I32_MAX - 1, I32_MAX - 1}};
}

enum FirstLast { FL_BEGINNING = 0, FL_IN_BOUNDS = 1, FL_ENDED = 2 };

// raycast.rs:301-310
struct RaycastStep {
    int32_t cube[3];
    int face;
    double t_distance;
    double t_max[3];
};

// raycast.rs:63-148 (Raycaster + State + Parameters flattened into one struct)
struct Raycaster {
    // Parameters (raycast.rs:126-148)
    double origin[3];
    double dir[3];
    int32_t step[3];
    double t_delta[3];
    // State (raycast.rs:99-121)
    Aab bounds;
    double t_max[3];
    int32_t cube[3];
    int last_face;
    double last_t_distance;
    // Raycaster (raycast.rs:63-75)
    int first_last;
    bool include_exit;

    void init(const double origin[3], const double dir[3]);  // Raycaster::new, raycast.rs:196-202
    void within(const Aab &b, bool include_exit);             // raycast.rs:223-230
    bool next(RaycastStep *out);                              // raycast.rs:239-284

  private:
    void set_empty();                      // State::EMPTY, raycast.rs:502-509
    void current(RaycastStep *out) const;  // raycast.rs:548-557
    bool valid_for_stepping() const;       // raycast.rs:563-570
    bool do_step();                        // raycast.rs:577-626
    void fast_forward();                   // raycast.rs:632-704
    void oob(bool *enter, bool *exit_) const;  // raycast.rs:711-728
};

double scale_to_integer_step(double s, double ds);  // raycast.rs:797-819
int32_t signum_101(double x);                        // raycast.rs:782-788
bool cube_containing(const double p[3], int32_t out[3]);  // math/cube.rs:97-119
// RaycastStep::intersection_point, raycast.rs:409-439
void intersection_point(const RaycastStep &s, const double origin[3], const double dir[3], double out[3]);

}  // namespace orc

extern "C" {

// ---- DDA-level entry points (for the raycast/tests.rs KATs) ---------------------------------
double orc_scale_to_integer_step(double s, double ds);
// Runs Raycaster::new(origin, dir)[.within(bounds, include_exit)] and writes up to max_steps
// steps: cube xyz (int32 x3), face (int32), t_distance (double), intersection point (double x3).
// Returns the number of steps produced (the iterator is drained up to max_steps).
int orc_raycast(const double origin[3], const double dir[3], const int32_t *bounds_lo_hi_or_null,
                int include_exit, int max_steps, int32_t *out_cube_face /*4 per step*/,
                double *out_t /*1 per step*/, double *out_point /*3 per step*/);
// recursive_raycast (raycast.rs:458-476) from the n-th (0-based) step of the outer unbounded cast.
int orc_recursive_raycast(const double origin[3], const double dir[3], int nth, int resolution,
                          const int32_t bounds_lo_hi[6], int max_steps, double out_sub_ray[6],
                          int32_t *out_cube_face, double *out_t);

// ---- colour-level entry points -----------------------------------------------------------------
// apply_transmittance (raytracer_components.rs:215-258): out = rgba[4], emission_coeff
void orc_apply_transmittance(const float rgba[4], float thickness, float out[5]);
float orc_packed_light_lut(int v);           // light/data.rs:222 scalar_out
int orc_packed_light_scalar_in(float v);     // light/data.rs:213-217
void orc_to_srgb8(const float colorbuf[4], float exposure, int tone_mapping, float maximum_intensity,
                  uint8_t out[4]);           // encoder, renderer.rs:287-291

// ---- scene-level entry points -------------------------------------------------------------------
typedef struct orc_scene orc_scene;
orc_scene *orc_scene_create(const aicb_scene_desc *);
void orc_scene_destroy(orc_scene *);

// TraceStep / DepthStep streams for the surface.rs KATs. kind: 0 EnterSurface/Span, 1 Invisible,
// 2 EnterBlock.  rec[i] = {kind, t_distance, exit_t (depth mode, else NaN), ip.xyz, cube.xyz,
// voxel.xyz, resolution, normal, rgba[4]} as 18 doubles.
int orc_surface_steps(const orc_scene *, const double origin_dir[6], int depth_iter, int max_steps,
                      double *rec /*18 per step*/);

// accum_mode: 0 = ColorBuf (passive depth/hit observers), 1 = CharacterBuf-like first-hit
// accumulator (text.rs:52-123), 2 = DepthBuf alone (accum.rs:254-311).
// out_text (mode 1): per ray an int32: -2 '.', -1 ' ', -3 'X' (Incomplete), >=0 block index hit.
int orc_trace_rays(const orc_scene *, const double (*origin_dir)[6], size_t n, const aicb_options *,
                   int accum_mode, float (*out_colorbuf)[4], double *depth_or_null,
                   aicb_hit *hit_or_null, uint32_t *steps_or_null, int32_t *text_or_null);

// Full image (== RtRenderer::draw / draw_rgba with the Rayon dispatch replaced by std::thread
// rows).  Any output pointer may be NULL.  Returns cubes_traced summed.
uint64_t orc_render(const orc_scene *, const aicb_camera *, const aicb_options *,
                    const aicb_shard *shard_or_null, int accum_mode, int n_threads,
                    uint8_t (*out_srgb8)[4], float (*out_colorbuf)[4], double *depth_or_null,
                    aicb_hit *hit_or_null, uint32_t *steps_or_null, int32_t *text_or_null);
// As orc_render but only rows [row_begin,row_end) (bounded CPU-baseline samples); outputs are
// indexed from row_begin.
uint64_t orc_render_rows(const orc_scene *, const aicb_camera *, const aicb_options *, uint32_t row_begin,
                         uint32_t row_end, int n_threads, uint8_t (*out_srgb8)[4],
                         float (*out_colorbuf)[4]);
// Arbitrary rows (bounded CPU-baseline samples spread over the frame); outputs packed in list order.
uint64_t orc_render_rowlist(const orc_scene *, const aicb_camera *, const aicb_options *, const uint32_t *rows,
                            size_t n_rows, int n_threads, uint8_t (*out_srgb8)[4], float (*out_colorbuf)[4]);
void orc_pixel_ray(const aicb_camera *, uint32_t x, uint32_t y, int sample /* -1 centre, 0..3 AA */,
                   double out_origin_dir[6]);
int orc_hardware_threads(void);
// BlockSky of a scene as texels: faces NX..PZ then mean (sky.rs:54-82)
void orc_scene_block_sky(const orc_scene *, uint8_t out[7][4]);

// ---- light propagation (oracle/aic_light.cpp) -----------------------------------
typedef struct orc_light orc_light;
size_t orc_light_chart(float *weights_or_null, uint32_t *children_or_null);  // flat chart (generator.rs), root = 0
orc_light *orc_light_create(const aicb_scene_desc *);   // light field = desc.light or all NO_RAYS
void orc_light_destroy(orc_light *);
void orc_light_fast_evaluate(orc_light *);               // updater.rs:537-582
void orc_light_set_cubes(orc_light *, const int32_t (*cubes)[3], const uint16_t *ids, size_t n);  // Mutation::set
uint64_t orc_light_evaluate(orc_light *, uint8_t epsilon, uint64_t max_updates, uint8_t *max_diff_out);  // space.rs:1496-1527
void orc_light_compute(orc_light *, const int32_t (*cubes)[3], size_t n, uint8_t (*out)[4]);  // compute_light only
void orc_light_get(const orc_light *, uint8_t (*out)[4]);
void orc_light_set_field(orc_light *, const uint8_t (*in)[4]);
void orc_light_get_outside(const orc_light *, const int32_t cube[3], uint8_t out[4]);  // LightStorage::get incl. out of bounds
void orc_light_set_pop_order(orc_light *, int order);  // 0 lowest cube index first (default), 1 highest first
size_t orc_light_queue_len(const orc_light *);
int orc_light_queue_peek(const orc_light *);
uint64_t orc_light_node_visits(const orc_light *);
}
