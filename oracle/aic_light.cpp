// ORACLE — TEST INFRASTRUCTURE ONLY (see aic_oracle.hpp).
//
// CPU restatement of all-is-cubes' light propagation (secondary hot path):
//   space/light/chart/generator.rs  (ray pattern -> prefix tree -> flat chart)
//   space/light/updater.rs          (compute_light, walk_ray_tree, LightBuffer, apply_light_update,
//                                    fast_evaluate_light, modified_cube_needs_update)
//   space/light/queue.rs            (priority queue; order inside one priority is unspecified in the
//                                    reference — hash-table order — and is "lowest cube index first" here)
//   space.rs:1496-1527              (evaluate_light)
// The non-threaded variant of update_light_from_queue (updater.rs:262-279) is followed: pop one, compute,
// apply.  f32 arithmetic order is kept (FaceMap::sum = (nx+px)+(ny+py)+(nz+pz), face.rs:1053).
#include <cmath>
#include <cstring>
#include <array>
#include <atomic>
#include <thread>
#include <map>
#include <set>
#include <unordered_map>
#include <vector>

#include "aic_oracle.hpp"

namespace orc {

static inline float ps_clamped_l(float v) { return (v > 0.0f) ? v : 0.0f; }
static inline float ps_mul_l(float a, float b) {
    float v = a * b;
    return (v != v) ? 0.0f : v;
}
static inline float fm_sum(const float w[6]) { return (w[0] + w[3]) + (w[1] + w[4]) + (w[2] + w[5]); }  // NX..PZ = 0..5

// ---------------------------------------------------------------------------------------------
// chart (generator.rs:49-215)
// ---------------------------------------------------------------------------------------------
struct FlatNode {
    float weight[6];
    uint32_t children[6];  // 0 = none
};

struct TreeNode {
    int8_t cube[3];
    int children[6];  // index into pool, -1 none
    float weight[6];
};

static std::vector<FlatNode> build_chart() {
    std::vector<TreeNode> pool;
    pool.push_back(TreeNode{{0, 0, 0}, {-1, -1, -1, -1, -1, -1}, {0, 0, 0, 0, 0, 0}});
    const int R = 5;
    for (int x = -R; x <= R; x++)
        for (int y = -R; y <= R; y++)
            for (int z = -R; z <= R; z++) {
                if (!(std::abs(x) == R || std::abs(y) == R || std::abs(z) == R)) continue;
                // Vector3D::<f32>::normalize = v / v.length()
                float fx = (float)x, fy = (float)y, fz = (float)z;
                float len = std::sqrt(fx * fx + fy * fy + fz * fz);
                float d[3] = {fx / len, fy / len, fz / len};
                float cos6[6];
                for (int f = 0; f < 6; f++) {
                    // unit_vector.dot(direction).max(0): x*dx + y*dy + z*dz with a unit axis vector
                    float u[3] = {0, 0, 0};
                    u[f % 3] = (f < 3) ? -1.0f : 1.0f;
                    float dot = u[0] * d[0] + u[1] * d[1] + u[2] * d[2];
                    cos6[f] = std::fmax(dot, 0.0f);
                }
                // ray_to_steps (generator.rs:100-113)
                double o[3] = {0.5, 0.5, 0.5};
                double dd[3] = {(double)d[0], (double)d[1], (double)d[2]};
                Raycaster rc;
                rc.init(o, dd);
                RaycastStep st;
                int cur = 0;
                bool first = true;
                while (rc.next(&st)) {
                    if (!(st.t_distance <= 127.0)) break;
                    // root weight accumulates too (insert adds weight at every node on the path, incl. root)
                    if (first) {
                        for (int f = 0; f < 6; f++) pool[0].weight[f] += cos6[f];
                        first = false;
                        continue;  // skip first step: it is always the root
                    }
                    int8_t rel[3] = {(int8_t)st.cube[0], (int8_t)st.cube[1], (int8_t)st.cube[2]};
                    // Face::from_adjacency(self.cube, next.cube)
                    int dir = -1;
                    for (int a = 0; a < 3; a++) {
                        int diff = (int)rel[a] - (int)pool[cur].cube[a];
                        if (diff == 1) dir = 3 + a;
                        else if (diff == -1) dir = a;
                    }
                    int child = pool[cur].children[dir];
                    if (child < 0) {
                        child = (int)pool.size();
                        pool[cur].children[dir] = child;
                        pool.push_back(TreeNode{{rel[0], rel[1], rel[2]}, {-1, -1, -1, -1, -1, -1}, {0, 0, 0, 0, 0, 0}});
                    }
                    cur = child;
                    for (int f = 0; f < 6; f++) pool[cur].weight[f] += cos6[f];
                }
            }
    // tree_to_flat: any numbering is equivalent for the walk; use pool order (root = 0).
    std::vector<FlatNode> flat(pool.size());
    for (size_t i = 0; i < pool.size(); i++) {
        for (int f = 0; f < 6; f++) {
            flat[i].weight[f] = pool[i].weight[f];
            flat[i].children[f] = pool[i].children[f] < 0 ? 0u : (uint32_t)pool[i].children[f];
        }
    }
    return flat;
}

static const std::vector<FlatNode> &chart() {
    static const std::vector<FlatNode> c = build_chart();
    return c;
}

// ---------------------------------------------------------------------------------------------
// light state
// ---------------------------------------------------------------------------------------------
struct LBlock {  // the EvaluatedBlock members light reads (evaluated.rs:189-272)
    bool opaque[6];
    bool all_opaque;
    bool visible;
    float face_color[7][4];  // Within, NX..PZ  (face7_color)
    float emission[3];
    bool has_emission;
};

struct PL {
    uint8_t r, g, b, s;
    bool operator==(const PL &o) const { return r == o.r && g == o.g && b == o.b && s == o.s; }
};
static const PL L_OPAQUE = {0, 0, 0, 128}, L_NO_RAYS = {0, 0, 0, 1}, L_UNINIT = {0, 0, 0, 0};

}  // namespace orc

struct orc_light {
    orc::Aab bounds;
    int32_t size[3];
    std::vector<uint16_t> ids;
    std::vector<orc::PL> light;
    std::vector<orc::LBlock> blocks;
    orc::PL sky_faces[6];
    int max_distance;
    // queue: by priority -> set of cube linear indices; by cube -> priority
    std::map<int, std::set<size_t>> by_priority;
    std::unordered_map<size_t, int> by_cube;
    std::atomic<uint64_t> node_visits{0};
    int pop_order = 0;  // 0: lowest cube index first, 1: highest first (the reference's order is unspecified)
};

namespace orc {

static inline float lut(uint8_t v) { return orc_packed_light_lut(v); }
static uint8_t scalar_in_l(float v) { return (uint8_t)orc_packed_light_scalar_in(v); }

static bool l_index(const orc_light &L, const int32_t c[3], size_t *idx) {
    uint32_t dx = (uint32_t)c[0] - (uint32_t)L.bounds.lo[0], dy = (uint32_t)c[1] - (uint32_t)L.bounds.lo[1],
             dz = (uint32_t)c[2] - (uint32_t)L.bounds.lo[2];
    if ((dx >= (uint32_t)L.size[0]) | (dy >= (uint32_t)L.size[1]) | (dz >= (uint32_t)L.size[2])) return false;
    *idx = ((size_t)dx * L.size[1] + dy) * L.size[2] + dz;
    return true;
}
static void l_cube_of(const orc_light &L, size_t idx, int32_t c[3]) {
    c[2] = (int32_t)(idx % L.size[2]) + L.bounds.lo[2];
    c[1] = (int32_t)((idx / L.size[2]) % L.size[1]) + L.bounds.lo[1];
    c[0] = (int32_t)(idx / ((size_t)L.size[2] * L.size[1])) + L.bounds.lo[0];
}

static const LBlock &air_block() {
    static LBlock b = [] {
        LBlock a;
        std::memset(&a, 0, sizeof a);
        return a;
    }();
    return b;
}
// UpdateCtx::get_evaluated (updater.rs:615-621): out of bounds = AIR
static const LBlock &get_evaluated(const orc_light &L, const int32_t c[3]) {
    size_t idx;
    if (l_index(L, c, &idx)) return L.blocks[L.ids[idx]];
    return air_block();
}
static inline bool opaque_for_light(const LBlock &b) { return b.all_opaque && !b.has_emission; }  // updater.rs:1031

// BlockSky::light_outside (sky.rs:113-147)
static PL light_outside_l(const orc_light &L, const int32_t c[3]) {
    int n_equal = 0, n_less = 0, which = 0;
    for (int a = 0; a < 3; a++) {
        int32_t beyond = L.bounds.lo[a] - 1;
        if (beyond == c[a]) { n_equal++; which = a; } else if (beyond < c[a]) n_less++;
        if (c[a] == L.bounds.hi[a]) { n_equal++; which = 3 + a; } else if (c[a] < L.bounds.hi[a]) n_less++;
    }
    if (n_less == 6) return L_UNINIT;
    if (n_equal == 1 && n_less == 5) return L.sky_faces[which];
    return L_NO_RAYS;
}
// LightStorage::get (updater.rs:585-595)
static PL light_get(const orc_light &L, const int32_t c[3]) {
    size_t idx;
    if (l_index(L, c, &idx)) return L.light[idx];
    return light_outside_l(L, c);
}

// ---- queue (queue.rs) ------------------------------------------------------------------------------
static void q_insert(orc_light &L, size_t idx, int prio) {
    auto it = L.by_cube.find(idx);
    if (it != L.by_cube.end()) {
        if (it->second >= prio) return;
        L.by_priority[it->second].erase(idx);
        if (L.by_priority[it->second].empty()) L.by_priority.erase(it->second);
        it->second = prio;
    } else {
        L.by_cube[idx] = prio;
    }
    L.by_priority[prio].insert(idx);
}
static void q_remove(orc_light &L, size_t idx) {
    auto it = L.by_cube.find(idx);
    if (it == L.by_cube.end()) return;
    L.by_priority[it->second].erase(idx);
    if (L.by_priority[it->second].empty()) L.by_priority.erase(it->second);
    L.by_cube.erase(it);
}
static int q_peek(const orc_light &L) { return L.by_priority.empty() ? 0 : L.by_priority.rbegin()->first; }
static bool q_pop(orc_light &L, size_t *idx) {
    if (L.by_priority.empty()) return false;
    auto it = std::prev(L.by_priority.end());
    auto pick = L.pop_order ? std::prev(it->second.end()) : it->second.begin();
    *idx = *pick;
    it->second.erase(pick);
    if (it->second.empty()) L.by_priority.erase(it);
    L.by_cube.erase(*idx);
    return true;
}
// light_needs_update (updater.rs:107-111)
static void light_needs_update(orc_light &L, const int32_t c[3], int prio) {
    size_t idx;
    if (l_index(L, c, &idx)) q_insert(L, idx, prio);
}

enum { PRIO_NEWLY_VISIBLE = 250, PRIO_UNINIT = 210, PRIO_ESTIMATED = 200 };
static inline int prio_from_difference(int d) { return d / 2 + 1; }

// ---- LightBuffer (updater.rs:694-944) ---------------------------------------------------------------
struct LightBuffer {
    float incoming[3] = {0, 0, 0};
    float total_weight = 0.0f;
    std::vector<size_t> deps_cubes;  // as linear indices, or SIZE_MAX for out-of-bounds cubes
    std::vector<int32_t> deps_xyz;
    double max_dist_sq;
    // orc_light_compute_by_chains: when set, what would be added to incoming / total_weight is appended here instead
    // ({x0, x1, x2, weight}; weight 0 for the terms that leave total_weight alone)
    std::vector<std::array<float, 4>> *rec = nullptr;
    uint64_t visits = 0;   // chart nodes entered by this cube's walk
};
struct RayState {
    float alpha;
    float dw[6];
};

static void add_weighted_light(LightBuffer &b, const float color[3], float weight) {  // updater.rs:926-929
    float k = ps_clamped_l(weight);
    if (b.rec) {
        b.rec->push_back({ps_mul_l(color[0], k), ps_mul_l(color[1], k), ps_mul_l(color[2], k), weight});
        return;
    }
    for (int i = 0; i < 3; i++) b.incoming[i] = b.incoming[i] + ps_mul_l(color[i], k);
    b.total_weight += weight;
}

static void end_of_ray(const orc_light &L, LightBuffer &b, const RayState &rs, float bundle_weight, const float cw[6]) {
    if (bundle_weight > 0.0f) {  // updater.rs:889-924
        float terms[6][3];
        for (int f = 0; f < 6; f++) {
            PL p = L.sky_faces[f];
            float v[3] = {lut(p.r), lut(p.g), lut(p.b)};
            float k = ps_clamped_l(cw[f]);
            for (int i = 0; i < 3; i++) terms[f][i] = ps_mul_l(v[i], k);
        }
        float sky[3];
        float recip = 1.0f / fm_sum(cw);
        float kr = ps_clamped_l(recip);
        for (int i = 0; i < 3; i++) {
            float s = (terms[0][i] + terms[3][i]) + (terms[1][i] + terms[4][i]) + (terms[2][i] + terms[5][i]);
            sky[i] = ps_mul_l(s, kr);
        }
        float ka = ps_clamped_l(rs.alpha);
        float c[3] = {ps_mul_l(sky[0], ka), ps_mul_l(sky[1], ka), ps_mul_l(sky[2], ka)};
        add_weighted_light(b, c, bundle_weight);
    }
}

static void push_dep(const orc_light &L, LightBuffer &b, const int32_t c[3], bool dedupe_last) {
    if (dedupe_last && b.deps_xyz.size() >= 3) {
        size_t n = b.deps_xyz.size();
        if (b.deps_xyz[n - 3] == c[0] && b.deps_xyz[n - 2] == c[1] && b.deps_xyz[n - 1] == c[2]) return;
    }
    b.deps_xyz.push_back(c[0]);
    b.deps_xyz.push_back(c[1]);
    b.deps_xyz.push_back(c[2]);
    (void)L;
}

// LightBuffer::traverse (updater.rs:760-884)
static void traverse(const orc_light &L, LightBuffer &b, RayState &rs, const int32_t cube[3], int face7, const LBlock &ev,
                     bool *have_ahead, PL *ahead, bool have_behind, PL behind, const float cw[6]) {
    if (!ev.visible) return;
    bool hit_opaque_face = (face7 == 0) ? ev.all_opaque : ev.opaque[face7 - 1];
    if (hit_opaque_face && face7 == 0) {
        for (int f = 0; f < 6; f++) rs.dw[f] = 0.0f;
        rs.alpha = 0.0f;
        return;
    }
    // face7_color(face).clamp()
    float col[4];
    for (int i = 0; i < 3; i++) col[i] = ev.face_color[face7][i] > 1.0f ? 1.0f : ev.face_color[face7][i];
    col[3] = ev.face_color[face7][3];
    const float hit_alpha = col[3];
    float wprod[6];
    for (int f = 0; f < 6; f++) wprod[f] = rs.dw[f] * cw[f];
    if (hit_alpha > 0.0f && face7 != 0) {
        int32_t lc[3] = {cube[0], cube[1], cube[2]};  // hit.adjacent(): the cube the ray came from
        int ax = (face7 - 1) % 3;
        lc[ax] += (face7 >= 4) ? 1 : -1;
        PL stored = have_behind ? behind : light_get(L, lc);
        float sv[3] = {lut(stored.r), lut(stored.g), lut(stored.b)};
        float lf[3];
        for (int i = 0; i < 3; i++) lf[i] = ev.emission[i] + ps_mul_l(ps_mul_l(col[i], sv[i]), hit_alpha);  // emission + reflect
        float ka = ps_clamped_l(rs.alpha), kw = ps_clamped_l(fm_sum(wprod));
        if (b.rec) b.rec->push_back({ps_mul_l(ps_mul_l(lf[0], ka), kw), ps_mul_l(ps_mul_l(lf[1], ka), kw), ps_mul_l(ps_mul_l(lf[2], ka), kw), 0.0f});
        else for (int i = 0; i < 3; i++) b.incoming[i] = b.incoming[i] + ps_mul_l(ps_mul_l(lf[i], ka), kw);
        push_dep(L, b, lc, true);
        if (hit_opaque_face) rs.alpha = 0.0f;
        else rs.alpha *= 1.0f - hit_alpha;
    }
    if (hit_alpha < 1.0f) {
        float sv[3] = {0, 0, 0};
        if (face7 != 0) {
            if (!*have_ahead) {
                *ahead = light_get(L, cube);
                *have_ahead = true;
            }
            sv[0] = lut(ahead->r); sv[1] = lut(ahead->g); sv[2] = lut(ahead->b);
        }
        float kh = ps_clamped_l(hit_alpha);
        float lt[3];
        for (int i = 0; i < 3; i++) lt[i] = ev.emission[i] + ps_mul_l(sv[i], kh);
        float ka = ps_clamped_l(rs.alpha), kw = ps_clamped_l(fm_sum(wprod));
        if (b.rec) b.rec->push_back({ps_mul_l(ps_mul_l(lt[0], ka), kw), ps_mul_l(ps_mul_l(lt[1], ka), kw), ps_mul_l(ps_mul_l(lt[2], ka), kw), 0.0f});
        else for (int i = 0; i < 3; i++) b.incoming[i] = b.incoming[i] + ps_mul_l(ps_mul_l(lt[i], ka), kw);
        push_dep(L, b, cube, false);
        rs.alpha *= 1.0f - hit_alpha;
    }
}

// walk_ray_tree (updater.rs:427-529)
static float walk(orc_light &L, LightBuffer &b, const int32_t origin[3], const int32_t cube[3], int face7, uint32_t node_index,
                  bool have_prev, PL prev, RayState rs) {
    const FlatNode &node = chart()[node_index];
    b.visits++;
    float prod[6];
    for (int f = 0; f < 6; f++) prod[f] = node.weight[f] * rs.dw[f];
    float bundle = fm_sum(prod);
    if (bundle <= 0.0f) return bundle;
    double dx = ((double)cube[0] + 0.5) - ((double)origin[0] + 0.5), dy = ((double)cube[1] + 0.5) - ((double)origin[1] + 0.5),
           dz = ((double)cube[2] + 0.5) - ((double)origin[2] + 0.5);
    double dist2 = dx * dx + dy * dy + dz * dz;
    if (dist2 > b.max_dist_sq) {
        end_of_ray(L, b, rs, bundle, node.weight);
        return bundle;
    }
    size_t idx;
    if (!l_index(L, cube, &idx)) {
        end_of_ray(L, b, rs, bundle, node.weight);
        return bundle;
    }
    bool have_ahead = false;
    PL ahead = L_UNINIT;
    traverse(L, b, rs, cube, face7, L.blocks[L.ids[idx]], &have_ahead, &ahead, have_prev, prev, node.weight);
    if (!(rs.alpha > 0.0f)) {
        end_of_ray(L, b, rs, bundle, node.weight);
        return bundle;
    }
    float child_sum = 0.0f;
    for (int f = 0; f < 6; f++) {
        if (node.children[f]) {
            int32_t nc[3] = {cube[0], cube[1], cube[2]};
            nc[f % 3] += (f < 3) ? -1 : 1;
            int opp = (f < 3) ? f + 3 : f - 3;
            child_sum += walk(L, b, origin, nc, opp + 1, node.children[f], have_ahead, ahead, rs);
        }
    }
    end_of_ray(L, b, rs, std::fmax(bundle - child_sum, 0.0f), node.weight);
    return bundle;
}

// compute_light (updater.rs:368-418) + finish (:932-944)
static PL compute_light(orc_light &L, const int32_t cube[3], LightBuffer *out_buf) {
    LightBuffer b;
    b.max_dist_sq = (double)L.max_distance * (double)L.max_distance;
    const LBlock &ev = get_evaluated(L, cube);
    bool origin_opaque = ev.all_opaque;
    if (origin_opaque) {
        if (!opaque_for_light(ev)) add_weighted_light(b, ev.emission, 1.0f);
    } else {
        RayState rs;
        rs.alpha = 1.0f;
        if (ev.visible) {
            for (int f = 0; f < 6; f++) rs.dw[f] = 1.0f;
        } else {  // directions_to_seek_light (updater.rs:669-690)
            for (int f = 0; f < 6; f++) {
                int opp = (f < 3) ? f + 3 : f - 3;
                int32_t nf[3] = {cube[0], cube[1], cube[2]}, no[3] = {cube[0], cube[1], cube[2]};
                nf[f % 3] += (f < 3) ? -1 : 1;
                no[opp % 3] += (opp < 3) ? -1 : 1;
                rs.dw[f] = (get_evaluated(L, no).visible || get_evaluated(L, nf).has_emission) ? 1.0f : 0.0f;
            }
        }
        walk(L, b, cube, cube, 0, 0, false, L_UNINIT, rs);
    }
    PL result;
    float scale = ps_clamped_l(1.0f / std::fmax(b.total_weight, 1.0f));
    if (b.total_weight > 0.0f) {
        result = PL{scalar_in_l(ps_mul_l(b.incoming[0], scale)), scalar_in_l(ps_mul_l(b.incoming[1], scale)),
                    scalar_in_l(ps_mul_l(b.incoming[2], scale)), 255};
    } else if (origin_opaque) {
        result = L_OPAQUE;
    } else {
        result = L_NO_RAYS;
    }
    L.node_visits.fetch_add(b.visits, std::memory_order_relaxed);
    if (out_buf) *out_buf = b;
    return result;
}

static int difference_priority(PL a, PL b) {  // data.rs:193-211
    auto ad = [](int x, int y) { return x > y ? x - y : y - x; };
    int d = std::max(std::max(ad(a.r, b.r), ad(a.g, b.g)), ad(a.b, b.b));
    if (a.s != b.s) d = std::min(255, d + 255 / 4);
    return d;
}

// apply_light_update (updater.rs:295-363)
static int apply_light_update(orc_light &L, const int32_t cube[3], PL nv, const LightBuffer &b) {
    size_t idx;
    if (!l_index(L, cube, &idx)) return 0;
    PL old = L.light[idx];
    int diff = difference_priority(nv, old);
    if (diff > 0) {
        L.light[idx] = nv;
        for (int f = 0; f < 6; f++) {
            int32_t nc[3] = {cube[0], cube[1], cube[2]};
            nc[f % 3] += (f < 3) ? -1 : 1;
            size_t nidx;
            if (!l_index(L, nc, &nidx)) continue;
            PL &nl = L.light[nidx];
            if (nl.s == 0) {  // Uninitialized
                if (nl == nv) continue;
                if (L.blocks[L.ids[nidx]].all_opaque) continue;
                float v[3] = {lut(nv.r), lut(nv.g), lut(nv.b)};
                nl = PL{scalar_in_l(v[0]), scalar_in_l(v[1]), scalar_in_l(v[2]), 0};  // PackedLight::guess
            }
        }
        if (diff > 1) {
            int prio = prio_from_difference(diff);
            for (size_t i = 0; i + 2 < b.deps_xyz.size(); i += 3) {
                int32_t dc[3] = {b.deps_xyz[i], b.deps_xyz[i + 1], b.deps_xyz[i + 2]};
                light_needs_update(L, dc, prio);
            }
        }
    }
    return diff;
}

}  // namespace orc

using namespace orc;

extern "C" {

size_t orc_light_chart(float *weights /*6 per node or NULL*/, uint32_t *children /*6 per node or NULL*/) {
    const auto &c = chart();
    if (weights)
        for (size_t i = 0; i < c.size(); i++) std::memcpy(weights + 6 * i, c[i].weight, sizeof c[i].weight);
    if (children)
        for (size_t i = 0; i < c.size(); i++) std::memcpy(children + 6 * i, c[i].children, sizeof c[i].children);
    return c.size();
}

orc_light *orc_light_create(const aicb_scene_desc *d) {
    orc_light *L = new orc_light();
    for (int a = 0; a < 3; a++) {
        L->bounds.lo[a] = d->bounds.lower[a];
        L->bounds.hi[a] = d->bounds.lower[a] + (int32_t)d->bounds.size[a];
        L->size[a] = (int32_t)d->bounds.size[a];
    }
    size_t vol = (size_t)L->size[0] * L->size[1] * L->size[2];
    L->ids.assign(d->block_ids, d->block_ids + vol);
    L->light.assign(vol, L_NO_RAYS);
    if (d->light)
        for (size_t i = 0; i < vol; i++) L->light[i] = PL{d->light[i][0], d->light[i][1], d->light[i][2], d->light[i][3]};
    L->blocks.resize(d->n_blocks);
    for (size_t i = 0; i < d->n_blocks; i++) {
        const aicb_block_desc &bd = d->blocks[i];
        LBlock &b = L->blocks[i];
        b.all_opaque = true;
        for (int f = 0; f < 6; f++) {
            b.opaque[f] = (bd.light_opaque_faces >> f) & 1;
            b.all_opaque = b.all_opaque && b.opaque[f];
            std::memcpy(b.face_color[f + 1], bd.light_face_colors[f], 16);
        }
        std::memcpy(b.face_color[0], bd.light_color, 16);
        std::memcpy(b.emission, bd.light_emission, 12);
        b.has_emission = !(b.emission[0] == 0.0f && b.emission[1] == 0.0f && b.emission[2] == 0.0f);
        b.visible = bd.light_visible != 0;
    }
    // BlockSky (sky.rs:54-82): reuse the raytracer oracle's construction through a throw-away scene
    {
        aicb_scene_desc tmp = *d;
        orc_scene *s = orc_scene_create(&tmp);
        uint8_t sk[7][4];
        orc_scene_block_sky(s, sk);
        for (int f = 0; f < 6; f++) L->sky_faces[f] = PL{sk[f][0], sk[f][1], sk[f][2], sk[f][3]};
        orc_scene_destroy(s);
    }
    L->max_distance = d->light_max_distance;
    return L;
}
void orc_light_destroy(orc_light *L) { delete L; }

// fast_evaluate_light (updater.rs:537-582)
void orc_light_fast_evaluate(orc_light *L) {
    L->by_priority.clear();
    L->by_cube.clear();
    if (L->max_distance == 0) return;
    for (int32_t x = L->bounds.lo[0]; x < L->bounds.hi[0]; x++)
        for (int32_t z = L->bounds.lo[2]; z < L->bounds.hi[2]; z++) {
            bool covered = false;
            for (int32_t y = L->bounds.hi[1] - 1; y >= L->bounds.lo[1]; y--) {
                int32_t c[3] = {x, y, z};
                size_t idx;
                l_index(*L, c, &idx);
                const LBlock &ev = L->blocks[L->ids[idx]];
                if (opaque_for_light(ev)) {
                    covered = true;
                    L->light[idx] = L_OPAQUE;
                } else {
                    bool any = ev.visible;
                    for (int f = 0; f < 6 && !any; f++) {
                        int32_t nc[3] = {x, y, z};
                        nc[f % 3] += (f < 3) ? -1 : 1;
                        any = get_evaluated(*L, nc).visible;
                    }
                    if (any) {
                        q_insert(*L, idx, PRIO_ESTIMATED);
                        L->light[idx] = covered ? L_UNINIT : L->sky_faces[4];  // in_direction(PY)
                    } else {
                        L->light[idx] = L_NO_RAYS;
                    }
                }
            }
        }
}

// Mutation::set -> side_effects_of_set -> modified_cube_needs_update (updater.rs:135-173)
void orc_light_set_cubes(orc_light *L, const int32_t (*cubes)[3], const uint16_t *ids, size_t n) {
    for (size_t i = 0; i < n; i++) {
        size_t idx;
        if (!l_index(*L, cubes[i], &idx)) continue;
        if (L->ids[idx] == ids[i]) continue;  // setting the same block is a no-op in Mutation::set
        L->ids[idx] = ids[i];
        if (L->max_distance == 0) continue;
        const LBlock &ev = L->blocks[ids[i]];
        if (opaque_for_light(ev)) {
            L->light[idx] = L_OPAQUE;
            q_remove(*L, idx);
        } else {
            light_needs_update(*L, cubes[i], PRIO_NEWLY_VISIBLE);
        }
        for (int f = 0; f < 6; f++) {
            int32_t nc[3] = {cubes[i][0], cubes[i][1], cubes[i][2]};
            nc[f % 3] += (f < 3) ? -1 : 1;
            int opp = (f < 3) ? f + 3 : f - 3;
            if (!get_evaluated(*L, nc).opaque[opp]) light_needs_update(*L, nc, PRIO_NEWLY_VISIBLE);
        }
    }
}

// evaluate_light (space.rs:1496-1527) without the wall-clock budget: run until the queue's highest
// priority is <= Priority::from_difference(epsilon).
uint64_t orc_light_evaluate(orc_light *L, uint8_t epsilon, uint64_t max_updates, uint8_t *max_diff_out) {
    uint64_t count = 0;
    int max_diff = 0;
    if (L->max_distance == 0) return 0;
    const int eps = prio_from_difference(epsilon);
    while (count < max_updates) {
        if (q_peek(*L) <= eps) break;
        size_t idx;
        if (!q_pop(*L, &idx)) break;
        int32_t c[3];
        l_cube_of(*L, idx, c);
        LightBuffer b;
        PL nv = compute_light(*L, c, &b);
        int d = apply_light_update(*L, c, nv, b);
        if (d > max_diff) max_diff = d;
        count++;
    }
    if (max_diff_out) *max_diff_out = (uint8_t)max_diff;
    return count;
}

// update_light_from_queue as the reference runs it with its `auto-threads` feature (updater.rs:211-252): pop up to 32
// cubes, compute their light in parallel from the same stored light, apply the results one after the other in pop
// order.  The result does not depend on the number of threads.  (orc_light_evaluate above is the non-threaded
// variant, :254-270, which pops and applies one cube at a time.)
uint64_t orc_light_evaluate_threaded(orc_light *L, uint8_t epsilon, uint64_t max_updates, int n_threads, uint8_t *max_diff_out) {
    uint64_t count = 0;
    int max_diff = 0;
    if (L->max_distance == 0) return 0;
    const int eps = prio_from_difference(epsilon);
    constexpr int BATCH = 32;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > BATCH) n_threads = BATCH;
    struct Item { int32_t c[3]; PL nv; LightBuffer b; };
    std::vector<Item> items(BATCH);
    // workers spin on a generation counter: a batch is ~32 x 50 us of work, too short for thread creation per batch
    std::atomic<int> generation{0}, next{0}, done{0}, n_items{0};
    std::atomic<bool> quit{false};
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_items.load(std::memory_order_acquire)) break;
            items[i].b = LightBuffer();
            items[i].nv = compute_light(*L, items[i].c, &items[i].b);
            done.fetch_add(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++)
        pool.emplace_back([&]() {
            int seen = 0;
            for (;;) {
                while (generation.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_relaxed)) std::this_thread::yield();
                if (quit.load(std::memory_order_relaxed)) return;
                seen = generation.load(std::memory_order_acquire);
                work();
            }
        });
    while (count < max_updates) {
        int n = 0;
        while (n < BATCH && count + n < max_updates) {
            if (q_peek(*L) <= eps) break;
            size_t idx;
            if (!q_pop(*L, &idx)) break;
            l_cube_of(*L, idx, items[n].c);
            n++;
        }
        if (n == 0) break;
        done.store(0, std::memory_order_relaxed);
        next.store(0, std::memory_order_relaxed);
        n_items.store(n, std::memory_order_release);
        generation.fetch_add(1, std::memory_order_release);
        work();
        while (done.load(std::memory_order_acquire) < n) std::this_thread::yield();
        for (int i = 0; i < n; i++) {
            int d = apply_light_update(*L, items[i].c, items[i].nv, items[i].b);
            if (d > max_diff) max_diff = d;
            count++;
        }
    }
    quit.store(true);
    generation.fetch_add(1, std::memory_order_release);
    for (std::thread &t : pool) t.join();
    if (max_diff_out) *max_diff_out = (uint8_t)max_diff;
    return count;
}

void orc_light_compute(orc_light *L, const int32_t (*cubes)[3], size_t n, uint8_t (*out)[4]) {
    for (size_t i = 0; i < n; i++) {
        PL p = compute_light(*L, cubes[i], nullptr);
        out[i][0] = p.r; out[i][1] = p.g; out[i][2] = p.b; out[i][3] = p.s;
    }
}

// compute_light evaluated the way the CUDA chain walk does (all-is-cubes_b200/csrc/light_kernel.cuh:
// compute_light_chains), with the chain tables of the PRODUCT library handed in (aicb_light_chart_chains): every chain
// of the chart is walked on its own — traverse / end_of_ray per node exactly as walk() above, but the terms are
// recorded per chain instead of added —, chains in breadth-first order (a child starts from what its parent left at the
// branching node), and the terms are then added in the Euler tour of the chain tree.  tests/test_light_chains.py
// asserts that this gives the bits of the recursive walk: the argument the kernel's bit-exactness rests on (a chain's
// nodes share their weights, so `bundle - children` is exactly 0 except at chain ends; depth-first order of the
// terms = the Euler tour), checked on the CPU.
void orc_light_compute_by_chains(orc_light *Lp, const int32_t (*cubes)[3], size_t n, uint8_t (*out)[4], float (*raw_or_null)[4],
                                 const uint32_t *preorder, const uint32_t (*chains)[6], uint32_t n_chains, const uint16_t *euler,
                                 uint32_t n_euler) {
    orc_light &L = *Lp;
    const std::vector<FlatNode> &ch = chart();
    // cube offset and entry face of every node of the flat chart
    static std::vector<std::array<int32_t, 4>> place;
    if (place.empty()) {
        place.assign(ch.size(), {0, 0, 0, 0});
        std::vector<uint32_t> stack{0};
        while (!stack.empty()) {
            const uint32_t i = stack.back();
            stack.pop_back();
            for (int f = 0; f < 6; f++)
                if (ch[i].children[f]) {
                    std::array<int32_t, 4> p = place[i];
                    p[f % 3] += (f < 3) ? -1 : 1;
                    p[3] = ((f < 3) ? f + 3 : f - 3) + 1;   // face7 through which the child cube is entered
                    place[ch[i].children[f]] = p;
                    stack.push_back(ch[i].children[f]);
                }
        }
    }
    struct BranchState { float alpha; bool have; PL ahead; };
    for (size_t q = 0; q < n; q++) {
        const int32_t *cube = cubes[q];
        LightBuffer b;
        b.max_dist_sq = (double)L.max_distance * (double)L.max_distance;
        const LBlock &ev = get_evaluated(L, cube);
        const bool origin_opaque = ev.all_opaque;
        if (origin_opaque) {
            if (!opaque_for_light(ev)) add_weighted_light(b, ev.emission, 1.0f);
        } else {
            float dw[6];
            if (ev.visible) {
                for (int f = 0; f < 6; f++) dw[f] = 1.0f;
            } else {
                for (int f = 0; f < 6; f++) {
                    int opp = (f < 3) ? f + 3 : f - 3;
                    int32_t nf[3] = {cube[0], cube[1], cube[2]}, no[3] = {cube[0], cube[1], cube[2]};
                    nf[f % 3] += (f < 3) ? -1 : 1;
                    no[opp % 3] += (opp < 3) ? -1 : 1;
                    dw[f] = (get_evaluated(L, no).visible || get_evaluated(L, nf).has_emission) ? 1.0f : 0.0f;
                }
            }
            std::vector<std::vector<std::array<float, 4>>> entry(n_chains), pop(n_chains);
            std::vector<char> ready(n_chains, 0);
            std::vector<BranchState> branch(n_chains, BranchState{0.0f, false, L_UNINIT});
            ready[0] = 1;
            for (uint32_t c = 0; c < n_chains; c++) {
                if (!ready[c]) continue;
                const uint32_t first = chains[c][0], length = chains[c][1], kids = chains[c][2], first_child = chains[c][3];
                const uint32_t pb = chains[c][4], br = chains[c][5];
                RayState rs;
                for (int f = 0; f < 6; f++) rs.dw[f] = dw[f];
                bool have_prev = false;
                PL prev = L_UNINIT;
                if (pb == 0xffffu) rs.alpha = 1.0f;
                else { rs.alpha = branch[pb].alpha; have_prev = branch[pb].have; prev = branch[pb].ahead; }
                b.rec = &entry[c];
                bool alive = true;
                float bundle = 0.0f;
                uint32_t last_node = 0;
                for (uint32_t k = 0; k < length && alive; k++) {
                    const uint32_t node_index = preorder[first + k];
                    const FlatNode &node = ch[node_index];
                    last_node = node_index;
                    b.visits++;
                    float prod[6];
                    for (int f = 0; f < 6; f++) prod[f] = node.weight[f] * rs.dw[f];
                    bundle = fm_sum(prod);
                    if (bundle <= 0.0f) { alive = false; break; }
                    const std::array<int32_t, 4> &pl = place[node_index];
                    const int32_t nc[3] = {cube[0] + pl[0], cube[1] + pl[1], cube[2] + pl[2]};
                    const double dx = (double)pl[0], dy = (double)pl[1], dz = (double)pl[2];
                    size_t idx;
                    if (dx * dx + dy * dy + dz * dz > b.max_dist_sq || !l_index(L, nc, &idx)) {
                        end_of_ray(L, b, rs, bundle, node.weight);
                        alive = false;
                        break;
                    }
                    bool have_ahead = false;
                    PL ahead = L_UNINIT;
                    traverse(L, b, rs, nc, pl[3], L.blocks[L.ids[idx]], &have_ahead, &ahead, have_prev, prev, node.weight);
                    if (!(rs.alpha > 0.0f)) {
                        end_of_ray(L, b, rs, bundle, node.weight);
                        alive = false;
                        break;
                    }
                    have_prev = have_ahead;
                    prev = ahead;
                }
                if (!alive) continue;
                // alive at the chain's last node: the children start from here; the rest of the bundle ends here
                float child_sum = 0.0f;
                for (uint32_t j = 0; j < kids; j++) {
                    const FlatNode &cn = ch[preorder[chains[first_child + j][0]]];
                    float prod[6];
                    for (int f = 0; f < 6; f++) prod[f] = cn.weight[f] * rs.dw[f];
                    child_sum += fm_sum(prod);
                    ready[first_child + j] = 1;
                }
                if (kids) branch[br] = BranchState{rs.alpha, have_prev, prev};
                b.rec = &pop[c];
                end_of_ray(L, b, rs, std::fmax(bundle - child_sum, 0.0f), ch[last_node].weight);
            }
            b.rec = nullptr;
            for (uint32_t p = 0; p < n_euler; p++) {
                const uint32_t c = euler[p] & 0x7fffu;
                const std::vector<std::array<float, 4>> &terms = (euler[p] & 0x8000u) ? pop[c] : entry[c];
                for (const std::array<float, 4> &t : terms) {
                    for (int i = 0; i < 3; i++) b.incoming[i] = b.incoming[i] + t[i];
                    b.total_weight += t[3];
                }
            }
        }
        L.node_visits.fetch_add(b.visits, std::memory_order_relaxed);
        PL result;
        float scale = ps_clamped_l(1.0f / std::fmax(b.total_weight, 1.0f));
        if (b.total_weight > 0.0f)
            result = PL{scalar_in_l(ps_mul_l(b.incoming[0], scale)), scalar_in_l(ps_mul_l(b.incoming[1], scale)),
                        scalar_in_l(ps_mul_l(b.incoming[2], scale)), 255};
        else if (origin_opaque) result = L_OPAQUE;
        else result = L_NO_RAYS;
        out[q][0] = result.r; out[q][1] = result.g; out[q][2] = result.b; out[q][3] = result.s;
        if (raw_or_null) {
            raw_or_null[q][0] = b.incoming[0]; raw_or_null[q][1] = b.incoming[1]; raw_or_null[q][2] = b.incoming[2];
            raw_or_null[q][3] = b.total_weight;
        }
    }
}

// compute_light with the unquantised accumulators (incoming_light, total_rays) beside the packed result
void orc_light_compute_raw(orc_light *L, const int32_t (*cubes)[3], size_t n, uint8_t (*out)[4], float (*raw)[4]) {
    for (size_t i = 0; i < n; i++) {
        LightBuffer b;
        PL p = compute_light(*L, cubes[i], &b);
        out[i][0] = p.r; out[i][1] = p.g; out[i][2] = p.b; out[i][3] = p.s;
        raw[i][0] = b.incoming[0]; raw[i][1] = b.incoming[1]; raw[i][2] = b.incoming[2]; raw[i][3] = b.total_weight;
    }
}

void orc_light_get(const orc_light *L, uint8_t (*out)[4]) {
    for (size_t i = 0; i < L->light.size(); i++) {
        out[i][0] = L->light[i].r; out[i][1] = L->light[i].g; out[i][2] = L->light[i].b; out[i][3] = L->light[i].s;
    }
}
void orc_light_set_field(orc_light *L, const uint8_t (*in)[4]) {
    for (size_t i = 0; i < L->light.size(); i++) L->light[i] = PL{in[i][0], in[i][1], in[i][2], in[i][3]};
}
void orc_light_get_outside(const orc_light *L, const int32_t c[3], uint8_t out[4]) {
    PL p = light_get(*L, c);
    out[0] = p.r; out[1] = p.g; out[2] = p.b; out[3] = p.s;
}
void orc_light_set_pop_order(orc_light *L, int order) { L->pop_order = order; }
size_t orc_light_queue_len(const orc_light *L) { return L->by_cube.size(); }
int orc_light_queue_peek(const orc_light *L) { return q_peek(*L); }
uint64_t orc_light_node_visits(const orc_light *L) { return L->node_visits.load(); }
}
