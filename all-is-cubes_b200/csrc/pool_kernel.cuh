// aicb200 — the marching kernel with a per-CTA ray pool (second scheduler of SpaceRaytracer::trace_ray's loop).
//
// Same arithmetic, same phases and same streams as trace_kernel (trace_kernel.cuh): what changes is who runs a ray.
// In trace_kernel a lane owns a ray from its first step to its last, so a warp marches with only the lanes whose
// ray is between two events (8.8 of 32 on the bench frame).  Here the complete state of POOL_RAYS rays lives in
// shared memory and a lane only borrows a ray for one phase:
//   * lanes that need work claim a ray that is ready to march (or a free slot, which they fill from the ray list);
//   * a ray that reaches an event (visible surface, block entry, span end, level exit, end of ray) is stored back
//     and marked EVENT_READY; the lane takes another marchable ray;
//   * when enough events wait, a warp claims 32 of them and runs the HEAVY phase with full lanes, then keeps
//     marching those rays.
// A slot is claimed with one shared-memory CAS; lane L only looks at slots L, L+32, ... so a claim needs no
// warp-level negotiation.  Results are identical to trace_kernel's by construction (per-ray code is shared
// verbatim or restated line by line); tests/test_gpu_parity.py compares the two frame for frame.
//
// Not in this scheduler: the AUX outputs (depth, hit position, per-pixel steps, device counters) — those renders
// use trace_kernel<.., AUX = true>.
#pragma once

#include "trace_kernel.cuh"

#ifdef __CUDACC__
namespace aicb {

#ifndef AICB_POOL_RAYS
#define AICB_POOL_RAYS 192
#endif
constexpr int POOL_RAYS = AICB_POOL_RAYS;   // rays resident per CTA: 6 per lane residue, shared by the 4 warps
constexpr int POOL_SLOTS_PER_LANE = POOL_RAYS / 32;
static_assert(POOL_RAYS % 32 == 0, "POOL_RAYS must be a multiple of the warp size");

enum PoolSlotState : int { PS_FREE = 0, PS_MARCH_READY = 1, PS_EVENT_READY = 2, PS_BUSY = 3 };

// per-ray doubles
enum : int { PD_TMX, PD_TMY, PD_TMZ, PD_LAST_T, PD_TDX, PD_TDY, PD_TDZ, PD_TSCALE, PD_EV_T,
             PD_SV_TMX, PD_SV_TMY, PD_SV_TMZ, PD_SV_LAST_T, PD_PEND_T, PD_PEND_IP0, PD_PEND_IP1, PD_PEND_IP2, PD_T_TO_ABS,
             POOL_ND };
// per-ray words
enum : int { PW_RX, PW_RY, PW_RZ, PW_FACE, PW_IDX, PW_FLAGS, PW_NXYZ, PW_T, PW_STEPS, PW_EV, PW_EV_CELL,
             PW_RES, PW_BLK0Y, PW_BLK0Z, PW_PALOFF,
             PW_SV_RX, PW_SV_RY, PW_SV_RZ, PW_SV_FACE, PW_SV_IDX, PW_SV_VALID,
             PW_PEND_PAL, PW_PEND_CX, PW_PEND_CY, PW_PEND_CZ, PW_PEND_PACKED, PW_PEND_RES,
             PW_T_TO_VIEW, PW_FIRST_HIT, PW_LAST_HIT, PW_TASK, PW_SKY,
             POOL_NW };
constexpr size_t POOL_SMEM_BYTES = (size_t)POOL_RAYS * (POOL_ND * 8 + POOL_NW * 4 + 4);

// PW_FLAGS: bits 0-1 sx+1, 2-3 sy+1, 4-5 sz+1, 6 valid, 7 inner, 8 need_advance, 9 have_last, 10-11 lane state
// PW_EV:    event kind | post << 8

template <bool VOLUMETRIC, bool WIDE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MIN_BLOCKS_PER_SM)
pool_kernel(const __grid_constant__ TraceParams P, uint32_t n_chunk_tasks) {
    (void)n_chunk_tasks;
    const DeviceScene &S = P.scene;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    extern __shared__ __align__(16) unsigned char pool_raw[];
    double *const pool_d = reinterpret_cast<double *>(pool_raw);
    uint32_t *const pool_w = reinterpret_cast<uint32_t *>(pool_d + (size_t)POOL_ND * POOL_RAYS);
    int *const pool_state = reinterpret_cast<int *>(pool_w + (size_t)POOL_NW * POOL_RAYS);
    __shared__ uint32_t s_bin_start[N_BINS + 1];
    __shared__ int s_live;          // rays currently in the pool
    __shared__ int s_exhausted;     // the ray list has no more rays for this CTA
    __shared__ int s_event_ready;   // rays waiting in PS_EVENT_READY (scheduling hint only)

    for (int i = threadIdx.x; i < POOL_RAYS; i += blockDim.x) pool_state[i] = PS_FREE;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int b = 0; b < N_BINS; b++) { s_bin_start[b] = acc; acc += P.bin_count[b]; }
        s_bin_start[N_BINS] = acc;
        s_live = 0;
        s_exhausted = 0;
        s_event_ready = 0;
    }
    __syncthreads();
    const uint32_t n_listed = s_bin_start[N_BINS];

#define PD(k) pool_d[(k) * POOL_RAYS + slot]
#define PW(k) pool_w[(k) * POOL_RAYS + slot]

    unsigned long long cubes_traced = 0;
    uint32_t hit_base = 0xffffffffu, hit_used = 0;   // this warp's block of the hit stream
    unsigned long long dbg_t0 = 0, dbg_passes = 0, dbg_rays = 0;
    if (P.debug_warp_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));

    // ---- the ray this lane currently runs (registers) ------------------------------------------------
    int slot = -1;
    int st = ST_IDLE;
    Ray r;      // t_delta and signs only
    Caster c;
    bool valid = false, inner = false, need_advance = false, have_last = false;
    double t_scale = 1.0;
    int nx = 0, ny = 0, nz = 0;
    uint32_t blk0y = 0, blk0z = 0, pal_off = 0;
    int res = 1;
    float T = 1.f;
    uint32_t steps = 0;
    int ev_kind = 0, ev_post = 0;
    double ev_t = 0.0;
    uint32_t ev_cell = 0;
    bool full_state = false;   // registers hold the HEAVY-only members too (res, block bounds, palette offset)

    const bool want_ip = P.lighting >= AICB_LIGHT_COARSE;
    const bool have_fog = (P.fog != AICB_FOG_NONE) && P.include_sky;
    const float fog_blend = (P.fog == AICB_FOG_ABRUPT) ? 1.0f : (P.fog == AICB_FOG_COMPROMISE ? 0.5f : 0.0f);

    auto count_stop = [&]() -> bool {  // count_step_should_stop (sr.rs:625-656)
        steps += 1;
        if (steps > 1000) return true;
        return T < (1.0f / 256.0f);
    };
    auto pop_level = [&]() {
        inner = false;
        t_scale = 1.0;
        c.tmx = PD(PD_SV_TMX); c.tmy = PD(PD_SV_TMY); c.tmz = PD(PD_SV_TMZ); c.last_t = PD(PD_SV_LAST_T);
        c.rx = (int)PW(PW_SV_RX); c.ry = (int)PW(PW_SV_RY); c.rz = (int)PW(PW_SV_RZ); c.face = (int)PW(PW_SV_FACE);
        c.idx = PW(PW_SV_IDX);
        valid = PW(PW_SV_VALID) != 0;
        nx = S.size[0]; ny = S.size[1]; nz = S.size[2];
        need_advance = true;
    };
    // what a DDA step and its classification need
    auto load_march_state = [&]() {
        c.tmx = PD(PD_TMX); c.tmy = PD(PD_TMY); c.tmz = PD(PD_TMZ); c.last_t = PD(PD_LAST_T);
        r.tdx = PD(PD_TDX); r.tdy = PD(PD_TDY); r.tdz = PD(PD_TDZ);
        t_scale = PD(PD_TSCALE);
        c.rx = (int)PW(PW_RX); c.ry = (int)PW(PW_RY); c.rz = (int)PW(PW_RZ); c.face = (int)PW(PW_FACE); c.idx = PW(PW_IDX);
        const uint32_t f = PW(PW_FLAGS);
        r.sx = (int)(f & 3u) - 1; r.sy = (int)((f >> 2) & 3u) - 1; r.sz = (int)((f >> 4) & 3u) - 1;
        valid = (f >> 6) & 1u; inner = (f >> 7) & 1u; need_advance = (f >> 8) & 1u; have_last = (f >> 9) & 1u;
        st = (int)((f >> 10) & 3u);
        if (inner) {
            const uint32_t n = PW(PW_NXYZ);
            nx = (int)(n & 255u); ny = (int)((n >> 8) & 255u); nz = (int)(n >> 16);
        } else {
            nx = S.size[0]; ny = S.size[1]; nz = S.size[2];
        }
        T = __uint_as_float(PW(PW_T));
        steps = PW(PW_STEPS);
        full_state = false;
    };
    // + what only the HEAVY phase needs
    auto load_event_state = [&]() {
        load_march_state();
        ev_t = PD(PD_EV_T);
        const uint32_t e = PW(PW_EV);
        ev_kind = (int)(e & 255u); ev_post = (int)(e >> 8);
        ev_cell = PW(PW_EV_CELL);
        res = (int)PW(PW_RES); blk0y = PW(PW_BLK0Y); blk0z = PW(PW_BLK0Z); pal_off = PW(PW_PALOFF);
        full_state = true;
    };
    auto store_state = [&]() {
        PD(PD_TMX) = c.tmx; PD(PD_TMY) = c.tmy; PD(PD_TMZ) = c.tmz; PD(PD_LAST_T) = c.last_t;
        PD(PD_TSCALE) = t_scale;
        PW(PW_RX) = (uint32_t)c.rx; PW(PW_RY) = (uint32_t)c.ry; PW(PW_RZ) = (uint32_t)c.rz; PW(PW_FACE) = (uint32_t)c.face;
        PW(PW_IDX) = c.idx;
        PW(PW_FLAGS) = (uint32_t)(r.sx + 1) | ((uint32_t)(r.sy + 1) << 2) | ((uint32_t)(r.sz + 1) << 4) | ((valid ? 1u : 0u) << 6) |
                       ((inner ? 1u : 0u) << 7) | ((need_advance ? 1u : 0u) << 8) | ((have_last ? 1u : 0u) << 9) | ((uint32_t)st << 10);
        PW(PW_T) = __float_as_uint(T);
        PW(PW_STEPS) = steps;
        PD(PD_EV_T) = ev_t;
        PW(PW_EV) = (uint32_t)ev_kind | ((uint32_t)ev_post << 8);
        PW(PW_EV_CELL) = ev_cell;
        if (full_state) {   // a marching-only borrower never changes these
            PW(PW_NXYZ) = (uint32_t)nx | ((uint32_t)ny << 8) | ((uint32_t)nz << 16);
            PW(PW_RES) = (uint32_t)res; PW(PW_BLK0Y) = blk0y; PW(PW_BLK0Z) = blk0z; PW(PW_PALOFF) = pal_off;
        }
    };
    auto load_ray = [&](Ray &rr) {   // origin / direction come back from the ray record (HBM / L2), they are rarely needed
        const RayRecord *rec = P.ray_records + PW(PW_TASK);
        const double2 a = __ldg(reinterpret_cast<const double2 *>(&rec->ox));
        const double2 b = __ldg(reinterpret_cast<const double2 *>(&rec->oz));
        const double2 d = __ldg(reinterpret_cast<const double2 *>(&rec->dy));
        rr = r;
        rr.ox = a.x; rr.oy = a.y; rr.oz = b.x; rr.dx = b.y; rr.dy = d.x; rr.dz = d.y;
        rr.half_over_len = __ldg(&rec->half_over_len);
    };
    auto store_pending = [&](const PendingSurface &sf) {
        PD(PD_PEND_T) = sf.t; PD(PD_PEND_IP0) = sf.ip[0]; PD(PD_PEND_IP1) = sf.ip[1]; PD(PD_PEND_IP2) = sf.ip[2];
        PW(PW_PEND_PAL) = sf.pal; PW(PW_PEND_CX) = (uint32_t)sf.cube[0]; PW(PW_PEND_CY) = (uint32_t)sf.cube[1];
        PW(PW_PEND_CZ) = (uint32_t)sf.cube[2]; PW(PW_PEND_PACKED) = sf.packed; PW(PW_PEND_RES) = (uint32_t)sf.res;
    };
    auto load_pending = [&](PendingSurface &sf) {
        sf.t = PD(PD_PEND_T); sf.ip[0] = PD(PD_PEND_IP0); sf.ip[1] = PD(PD_PEND_IP1); sf.ip[2] = PD(PD_PEND_IP2);
        sf.pal = PW(PW_PEND_PAL); sf.cube[0] = (int)PW(PW_PEND_CX); sf.cube[1] = (int)PW(PW_PEND_CY);
        sf.cube[2] = (int)PW(PW_PEND_CZ); sf.packed = PW(PW_PEND_PACKED); sf.res = (int)PW(PW_PEND_RES);
    };

    // Warp-cooperative claim: every lane with `need` gets a slot that is in state `want` as long as there are any.
    // The warp reads the state array 32 slots at a time; the first candidates (as many as lanes still need one) try
    // the CAS, and the i-th lane in need takes the i-th slot that was won.  The scan starts at a per-warp offset so
    // that the four warps do not fight over the same slots first.
    const unsigned lanemask_lt = (1u << lane) - 1u;
    auto claim = [&](int want, bool need) -> int {
        volatile int *vs = pool_state;
        int got = -1;
        unsigned need_mask = __ballot_sync(0xffffffffu, need);
        for (int j = 0; j < POOL_SLOTS_PER_LANE && need_mask; j++) {
            int jj = j + warp;
            if (jj >= POOL_SLOTS_PER_LANE) jj -= POOL_SLOTS_PER_LANE;
            const int s = lane + 32 * jj;
            const bool cand = vs[s] == want;
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (!m) continue;
            const bool won = cand && __popc(m & lanemask_lt) < __popc(need_mask) && atomicCAS(pool_state + s, want, PS_BUSY) == want;
            const unsigned m2 = __ballot_sync(0xffffffffu, won);
            if (need && got < 0) {
                const int my_rank = __popc(need_mask & lanemask_lt);
                if (my_rank < __popc(m2)) got = (int)__fns(m2, 0, my_rank + 1) + 32 * jj;
            }
            need_mask = __ballot_sync(0xffffffffu, need && got < 0);
        }
        return got;
    };
    // give the ray back to the pool in state `next` (after its state has been stored)
    auto publish = [&](int next) {
        __threadfence_block();
        atomicExch(pool_state + slot, next);
        slot = -1;
        st = ST_IDLE;
    };

    auto record_surface = [&](PendingSurface &sf, uint32_t cell_or_voxel, double t) {
        int cx, cy, cz;
        if (!inner) {
            const uint4 *bp = reinterpret_cast<const uint4 *>(S.blocks + cell_or_voxel);
            sf.pal = __ldg(bp + 1).y;
            cx = c.rx + S.lo[0]; cy = c.ry + S.lo[1]; cz = c.rz + S.lo[2];
            sf.packed = (uint32_t)c.face << 24;
            sf.res = 1;
        } else {
            sf.pal = pal_off + cell_or_voxel;
            cx = (int)PW(PW_SV_RX) + S.lo[0]; cy = (int)PW(PW_SV_RY) + S.lo[1]; cz = (int)PW(PW_SV_RZ) + S.lo[2];
            const int vx = c.rx + (int)(int16_t)(blk0y & 0xffff), vy = c.ry + (int)(int16_t)(blk0y >> 16),
                      vz = c.rz + (int)(int16_t)(blk0z & 0xffff);
            sf.packed = (uint32_t)vx | ((uint32_t)vy << 8) | ((uint32_t)vz << 16) | ((uint32_t)c.face << 24);
            sf.res = res;
        }
        sf.t = t;
        sf.cube[0] = cx; sf.cube[1] = cy; sf.cube[2] = cz;
        if (want_ip) {
            double ip[3];
            Ray rr;
            load_ray(rr);
            if (!inner) {
                intersection_point(c, rr, cx, cy, cz, rr.ox, rr.oy, rr.oz, ip);
            } else {
                const double fres = (double)res, anti = recip_pow2(res);
                const int vx = (int)(sf.packed & 255), vy = (int)((sf.packed >> 8) & 255), vz = (int)((sf.packed >> 16) & 255);
                intersection_point(c, rr, vx, vy, vz, (rr.ox - (double)cx) * fres, (rr.oy - (double)cy) * fres,
                                   (rr.oz - (double)cz) * fres, ip);
                ip[0] = ip[0] * anti + (double)cx;  // surface.rs:406-407
                ip[1] = ip[1] * anti + (double)cy;
                ip[2] = ip[2] * anti + (double)cz;
            }
            sf.ip[0] = ip[0]; sf.ip[1] = ip[1]; sf.ip[2] = ip[2];
        }
    };
    // One DDA step and the classification of the cube / voxel it lands on (as trace_kernel's march_step).
    auto march_step = [&]() {
        if (need_advance) {
            if (!valid) { ev_kind = EV_STUCK; st = ST_EVENT; return; }  // raycast.rs:245-249
            if (caster_step(c, r, nx, ny, nz)) {
                ev_kind = EV_EXIT; ev_t = c.last_t * t_scale; st = ST_EVENT;
                return;
            }
        }
        need_advance = true;
        uint32_t word;
        bool invisible, enter_block;
        if constexpr (WIDE) {
            if (!inner) {
                const uint32_t cell = __ldg((const uint32_t *)S.cells + c.idx);
                word = cell & 0xffffu;
                invisible = (cell >> 16) == KIND_INVISIBLE;
                enter_block = (cell >> 16) == KIND_RECURSIVE;
            } else {
                word = __ldg(S.bricks + c.idx);
                invisible = (word & 0x8000u) != 0;
                enter_block = false;
            }
        } else {
            const uint16_t *vol = inner ? S.bricks : (const uint16_t *)S.cells;
            const uint32_t w = __ldg(vol + c.idx);
            invisible = (w & 0x8000u) != 0;
            enter_block = !inner & ((w & 0x4000u) != 0);
            word = inner ? w : (w & 0x3fffu);
        }
        if (invisible) {
            if (VOLUMETRIC && have_last) {
                ev_kind = EV_INVISIBLE; ev_t = c.last_t * t_scale; ev_post = POST_CONTINUE;
                st = ST_EVENT;
                return;
            }
            if (count_stop()) st = ST_DONE;
            return;
        }
        ev_kind = enter_block ? EV_ENTER_BLOCK : EV_SURFACE;
        ev_t = c.last_t * t_scale;
        ev_cell = word;
        ev_post = POST_CONTINUE;
        st = ST_EVENT;
    };

    // take a marchable ray for every lane that has none: first one that waits in the pool, else a new one from the list
    auto acquire_marchable = [&]() {
        const int got = claim(PS_MARCH_READY, slot < 0);
        if (got >= 0) {
            slot = got;
            __threadfence_block();
            load_march_state();
            st = ST_MARCH;
        }
        // new rays
        const bool exhausted = *(volatile int *)&s_exhausted != 0;
        const int fs = claim(PS_FREE, slot < 0 && !exhausted);
        const unsigned m = __ballot_sync(0xffffffffu, fs >= 0);
        if (m) {
            const int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(P.task_counter, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (fs >= 0) {
                const uint32_t k = base + __popc(m & ((1u << lane) - 1u));
                if (k >= n_listed) {
                    atomicExch(pool_state + fs, PS_FREE);
                    s_exhausted = 1;
                } else {
                    slot = fs;
                    atomicAdd(&s_live, 1);
                    int b = 0;
                    while (k >= s_bin_start[b + 1]) b++;
                    const uint32_t task = __ldg(P.bin_list + (size_t)b * P.bin_stride + (k - s_bin_start[b]));
                    RayRecord rec;
                    {
                        const uint4 *src = reinterpret_cast<const uint4 *>(P.ray_records + task);
                        uint4 *dst = reinterpret_cast<uint4 *>(&rec);
#pragma unroll
                        for (int q = 0; q < 9; q++) dst[q] = __ldg(src + q);
                    }
                    r.tdx = rec.tdx; r.tdy = rec.tdy; r.tdz = rec.tdz;
                    r.sx = (int)((rec.flags >> 6) & 3u) - 1; r.sy = (int)((rec.flags >> 8) & 3u) - 1;
                    r.sz = (int)((rec.flags >> 10) & 3u) - 1;
                    c.tmx = rec.tmx; c.tmy = rec.tmy; c.tmz = rec.tmz; c.last_t = rec.last_t;
                    c.rx = rec.rx; c.ry = rec.ry; c.rz = rec.rz;
                    c.idx = rec.idx;
                    c.face = (int)(rec.flags & 7u);
                    valid = (rec.flags & 16u) != 0;
                    PD(PD_TDX) = rec.tdx; PD(PD_TDY) = rec.tdy; PD(PD_TDZ) = rec.tdz;
                    PD(PD_T_TO_ABS) = rec.t_to_abs;
                    PW(PW_T_TO_VIEW) = __float_as_uint(rec.t_to_view);
                    PW(PW_SKY) = (rec.flags >> 12) & 7u;
                    PW(PW_FIRST_HIT) = 0xffffffffu;
                    PW(PW_LAST_HIT) = 0xffffffffu;
                    PW(PW_TASK) = task;
                    T = 1.0f;
                    steps = 0;
                    have_last = false;
                    inner = false;
                    t_scale = 1.0;
                    need_advance = false;
                    nx = S.size[0]; ny = S.size[1]; nz = S.size[2];
                    res = 1; blk0y = blk0z = pal_off = 0;
                    ev_kind = 0; ev_post = 0; ev_t = 0.0; ev_cell = 0;
                    full_state = true;
                    st = ST_MARCH;
                }
            }
        }
    };

    constexpr int EVENT_BATCH_MIN = 32;   // run the HEAVY phase once this many events wait in the pool
    constexpr int SWAP_MIN = 8;           // lanes that finished marching before the warp stops to swap rays
    constexpr int EVENT_PRESSURE = POOL_RAYS / 2;

    for (;;) {
        dbg_passes++;
        // =========================== ACQUIRE ===============================================================
        {
            int evr = *(volatile int *)&s_event_ready;
            evr = __shfl_sync(0xffffffffu, evr, 0);
            bool want_events = evr >= EVENT_BATCH_MIN;
            for (int attempt = 0; attempt < 2; attempt++) {
                if (want_events) {
                    const int got = claim(PS_EVENT_READY, slot < 0);
                    if (got >= 0) {
                        slot = got;
                        __threadfence_block();
                        load_event_state();   // st comes back as ST_EVENT or ST_DONE
                    }
                    const int n = __popc(__ballot_sync(0xffffffffu, got >= 0));
                    if (n && lane == 0) atomicSub(&s_event_ready, n);
                } else {
                    acquire_marchable();
                }
                if (__any_sync(0xffffffffu, slot >= 0)) break;
                want_events = !want_events;   // nothing of the preferred kind: take the other
            }
            if (!__any_sync(0xffffffffu, slot >= 0)) {
                const int live = *(volatile int *)&s_live;
                const int ex = *(volatile int *)&s_exhausted;
                if (__shfl_sync(0xffffffffu, (ex != 0 && live == 0) ? 1 : 0, 0)) break;
                __nanosleep(100);
                continue;
            }
        }

        // =========================== HEAVY: events (as trace_kernel) =========================================
        if (__any_sync(0xffffffffu, slot >= 0 && (st == ST_EVENT || st == ST_DONE))) {
            // (0) leaving the level / an iterator that cannot step
            if (st == ST_EVENT && ev_kind == EV_STUCK) {   // ends without an exit step (raycast.rs:245-249)
                if (inner) { pop_level(); st = ST_MARCH; } else { st = ST_DONE; }
            }
            if (st == ST_EVENT && ev_kind == EV_EXIT) {
                // exit step: TraceStep::Invisible at this t (surface.rs:296-301, 388-393)
                if (VOLUMETRIC && have_last) {
                    ev_kind = EV_INVISIBLE; ev_post = inner ? POST_POP : POST_FINISH;
                } else if (count_stop() || !inner) {
                    st = ST_DONE;
                } else {
                    pop_level();
                    st = ST_MARCH;
                }
            }
            // (1) DepthIter + the Volumetric loop (surface.rs:460-490, sr.rs:185-203) / the Surface loop (sr.rs:206-225)
            bool do_shade = false;
            PendingSurface shade_sf;
            double span_exit = 0.0;
            if (st == ST_EVENT) {
                bool stop;
                if constexpr (VOLUMETRIC) {
                    do_shade = have_last;
                    load_pending(shade_sf);
                    span_exit = ev_t;
                    have_last = false;
                    stop = count_stop();
                } else {
                    stop = count_stop();
                    if (!stop && ev_kind == EV_SURFACE) {
                        record_surface(shade_sf, ev_cell, ev_t);
                        do_shade = true;
                    }
                }
                if (stop) {
                    do_shade = false;
                    st = ST_DONE;
                }
            }
            // (2) apply_transmittance, limit_alpha, invisibility test, fog amount, transmittance update
            bool emit = false;
            float h_ca = 0.f, h_coeff = 0.f, h_fa = -1.0f, h_tr = 1.0f;
            bool h_zeroed = false;
            if (do_shade) {
                const float4 col = __ldg(S.palette + 2 * (size_t)shade_sf.pal);
                const float4 emi = __ldg(S.palette + 2 * (size_t)shade_sf.pal + 1);
                float ca = col.w;
                float coeff = 1.0f;
                bool zeroed = false;
                if constexpr (VOLUMETRIC) {
                    const float thickness = fmaxf((float)((span_exit - shade_sf.t) * PD(PD_T_TO_ABS)), 0.0f);
                    if (thickness == 0.0f) {
                        if (col.w == 1.0f) { coeff = 1.0f; }
                        else { zeroed = true; ca = 0.0f; coeff = 0.0f; }
                    } else if (col.w == 1.0f) {
                        ca = 1.0f; coeff = 1.0f;        // 0^thickness == 0 exactly: alpha 1, (0-1)/(0-1) == 1
                    } else if (col.w == 0.0f) {
                        ca = 0.0f; coeff = thickness;   // 1^thickness == 1 exactly
                    } else {
                        const float unit_t = 1.0f - col.w;
                        const float depth_t = powf_exact(unit_t, thickness);
                        ca = zo_clamped(1.0f - depth_t);
                        const float k = (unit_t == 1.0f) ? thickness : (depth_t - 1.0f) / (unit_t - 1.0f);
                        coeff = fmaxf(k, 0.0f);
                    }
                }
                const float kc = ps_clamped(coeff);
                const float er = VOLUMETRIC ? ps_mul(emi.x, kc) : emi.x, eg = VOLUMETRIC ? ps_mul(emi.y, kc) : emi.y,
                            eb = VOLUMETRIC ? ps_mul(emi.z, kc) : emi.z;
                if (P.transparency == AICB_TRANSPARENCY_THRESHOLD) {  // limit_alpha (graphics_options.rs:496-507)
                    if (ca > P.threshold) { ca = 1.0f; } else { zeroed = true; ca = 0.0f; }
                }
                if (!(ca == 0.0f && er == 0.0f && eg == 0.0f && eb == 0.0f)) {
                    float tr = 1.0f - ca;
                    float fa = -1.0f;
                    if (have_fog) {  // distance_fog (sr.rs:745-768)
                        float rel = (float)shade_sf.t * __uint_as_float(PW(PW_T_TO_VIEW));
                        rel = rel < 0.0f ? 0.0f : (rel > 1.0f ? 1.0f : rel);
                        const float fog_exponential = 1.0f - expf_exact(-1.6f * rel);
                        const float fudged = fog_exponential / 0.79810348f;
                        const float p4 = (rel * rel) * (rel * rel);
                        fa = zo_clamped(fudged * (1.0f - fog_blend) + p4 * fog_blend);
                        tr = tr * (1.0f - fa);
                    }
                    emit = true;
                    h_ca = ca; h_coeff = coeff; h_fa = fa; h_tr = tr; h_zeroed = zeroed;
                }
            }
            // emit the hits of this pass (slots from the warp's block of the hit stream)
            {
                const unsigned em = __ballot_sync(0xffffffffu, emit);
                if (em) {
                    const uint32_t n_emit = (uint32_t)__popc(em);
                    if (hit_base == 0xffffffffu || hit_used + n_emit > HIT_BLOCK) {
                        if (hit_base != 0xffffffffu && hit_used + (uint32_t)lane < HIT_BLOCK)
                            P.hits[hit_base + hit_used + lane].pal = HIT_DEAD;   // fewer than 32 slots are left over
                        uint32_t nb = 0;
                        if (lane == 0) nb = atomicAdd(P.hit_counter, HIT_BLOCK);
                        nb = __shfl_sync(0xffffffffu, nb, 0);
                        if (nb >= P.hit_capacity) {  // (the capacity is a multiple of HIT_BLOCK)
                            if (lane == 0) *P.overflow_flag = 1u;  // the host re-runs the frame with a larger buffer
                            nb = 0xffffffffu;
                        }
                        hit_base = nb;
                        hit_used = 0;
                    }
                    if (emit) {
                        if (hit_base != 0xffffffffu) {
                            const uint32_t hslot = hit_base + hit_used + (uint32_t)__popc(em & ((1u << lane) - 1u));
                            HitRecord h;
                            if (want_ip) { h.ip[0] = shade_sf.ip[0]; h.ip[1] = shade_sf.ip[1]; h.ip[2] = shade_sf.ip[2]; }
                            else { h.ip[0] = h.ip[1] = h.ip[2] = 0.0; }
                            h.pal = shade_sf.pal;
                            h.cube[0] = shade_sf.cube[0]; h.cube[1] = shade_sf.cube[1]; h.cube[2] = shade_sf.cube[2];
                            h.T_before = T;
                            h.ca = h_ca;
                            h.coeff = h_coeff;
                            h.fa = h_fa;
                            h.flags = (shade_sf.packed >> 24) | (h_zeroed ? 8u : 0u) | (PW(PW_SKY) << 4);
                            h.next = 0xffffffffu;
                            const uint4 *src = reinterpret_cast<const uint4 *>(&h);
                            uint4 *dst = reinterpret_cast<uint4 *>(P.hits + hslot);
#pragma unroll
                            for (int q = 0; q < 4; q++) st_stream(dst + q, src[q]);
                            const uint32_t prev_hit = PW(PW_LAST_HIT);
                            if (prev_hit != 0xffffffffu) P.hits[prev_hit].next = hslot; else PW(PW_FIRST_HIT) = hslot;
                            PW(PW_LAST_HIT) = hslot;
                        }
                        T = T * h_tr;
                    }
                    if (hit_base != 0xffffffffu) hit_used += n_emit;
                }
            }
            // (3) Volumetric: the surface that raised this event becomes the pending one (surface.rs:467-476)
            if constexpr (VOLUMETRIC) {
                if (st == ST_EVENT && ev_kind == EV_SURFACE) {
                    PendingSurface pending;
                    record_surface(pending, ev_cell, ev_t);
                    store_pending(pending);
                    have_last = true;
                }
            }
            // (4b) the buffered DepthStep::EnterBlock is counted after the flushed span was traced
            if constexpr (VOLUMETRIC) {
                if (st == ST_EVENT && ev_kind == EV_ENTER_BLOCK) {
                    if (count_stop()) st = ST_DONE;
                }
            }
            // (5) recursive_raycast (raycast.rs:458-476) + TraceStep::EnterBlock (surface.rs:334-352)
            if (st == ST_EVENT && ev_kind == EV_ENTER_BLOCK) {
                const uint4 *bp = reinterpret_cast<const uint4 *>(S.blocks + ev_cell);
                const uint4 b0 = __ldg(bp);
                const uint4 b1 = __ldg(bp + 1);
                const int bres = (int)(b0.x >> 8);
                Level in;
                in.lox = (int16_t)(b0.y & 0xffff); in.loy = (int16_t)(b0.y >> 16); in.loz = (int16_t)(b0.z & 0xffff);
                in.nx = (int)(b0.z >> 16); in.ny = (int)(b0.w & 0xffff); in.nz = (int)(b0.w >> 16);
                in.base = b1.x;
                const double fres = (double)bres;
                const int cx = c.rx + S.lo[0], cy = c.ry + S.lo[1], cz = c.rz + S.lo[2];
                Caster ic;
                bool ivalid;
                Ray rr;
                load_ray(rr);
                if (caster_begin(ic, rr, (rr.ox - (double)cx) * fres, (rr.oy - (double)cy) * fres, (rr.oz - (double)cz) * fres, in,
                                 &ivalid)) {
                    PD(PD_SV_TMX) = c.tmx; PD(PD_SV_TMY) = c.tmy; PD(PD_SV_TMZ) = c.tmz; PD(PD_SV_LAST_T) = c.last_t;
                    PW(PW_SV_RX) = (uint32_t)c.rx; PW(PW_SV_RY) = (uint32_t)c.ry; PW(PW_SV_RZ) = (uint32_t)c.rz;
                    PW(PW_SV_FACE) = (uint32_t)c.face; PW(PW_SV_IDX) = c.idx; PW(PW_SV_VALID) = valid ? 1u : 0u;
                    c = ic;
                    valid = ivalid;
                    inner = true;
                    nx = in.nx; ny = in.ny; nz = in.nz;
                    blk0y = b0.y; blk0z = b0.z;
                    pal_off = b1.y;
                    res = bres;
                    t_scale = recip_pow2(bres);
                    need_advance = false;
                }
            }
            // (6) what the event's producer wanted next
            if (st == ST_EVENT) {
                if (ev_post == POST_POP) pop_level();
                st = (ev_post == POST_FINISH) ? ST_DONE : ST_MARCH;
            }
            // FINALIZE: hand the ray's result to the encode kernel, free the slot
            if (slot >= 0 && st == ST_DONE) {
                cubes_traced += steps;
                dbg_rays++;
                TaskOut o;
                o.first_hit = PW(PW_FIRST_HIT);
                o.T = T;
                o.steps = steps;
                o.flags = PW(PW_SKY);
                *reinterpret_cast<uint4 *>(P.task_out + PW(PW_TASK)) = *reinterpret_cast<const uint4 *>(&o);
                atomicSub(&s_live, 1);
                publish(PS_FREE);
            }
        }

        // =========================== MARCH ==================================================================
        {
            int iter = 0;
            bool pressure = false;
            for (;;) {
                if (!__any_sync(0xffffffffu, slot >= 0 && st == ST_MARCH)) break;
                if (slot >= 0 && st == ST_MARCH) march_step();
                if ((++iter & 3) == 0) {
                    const unsigned waiting = __ballot_sync(0xffffffffu, slot < 0 || st != ST_MARCH);
                    if (__popc(waiting) >= SWAP_MIN) {
                        // lanes whose ray reached an event hand it to the pool and take a marchable one
                        const bool fin = slot >= 0 && st != ST_MARCH;
                        if (fin) { store_state(); publish(PS_EVENT_READY); }
                        const int nf = __popc(__ballot_sync(0xffffffffu, fin));
                        if (nf && lane == 0) atomicAdd(&s_event_ready, nf);
                        acquire_marchable();
                        int evr = *(volatile int *)&s_event_ready;
                        evr = __shfl_sync(0xffffffffu, evr, 0);
                        if (evr >= EVENT_PRESSURE) { pressure = true; break; }
                    }
                }
            }
            // hand everything back: events wait for a HEAVY pass, rays still marching (pressure) for a lane
            const bool fin = slot >= 0 && st != ST_MARCH;
            const bool mid = slot >= 0 && st == ST_MARCH;
            if (fin) { store_state(); publish(PS_EVENT_READY); }
            if (mid) { store_state(); publish(PS_MARCH_READY); }
            const int nf = __popc(__ballot_sync(0xffffffffu, fin));
            if (nf && lane == 0) atomicAdd(&s_event_ready, nf);
            (void)pressure;
        }
    }
#undef PD
#undef PW

    if (P.debug_warp_times) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        for (int off = 16; off > 0; off >>= 1) dbg_rays += __shfl_down_sync(0xffffffffu, dbg_rays, off);
        if (lane == 0) {
            unsigned long long *d = P.debug_warp_times + 4 * (size_t)(blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5));
            d[0] = dbg_t0; d[1] = t1; d[2] = dbg_passes; d[3] = dbg_rays;
        }
    }
    // the unused rest of this warp's block of the hit stream
    if (hit_base != 0xffffffffu)
        for (uint32_t j = hit_used + lane; j < HIT_BLOCK; j += 32) P.hits[hit_base + j].pal = HIT_DEAD;
    // RaytraceInfo sum (renderer.rs:555): warp-reduce then one atomic per warp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cubes_traced += __shfl_down_sync(0xffffffffu, cubes_traced, off);
    if (lane == 0 && cubes_traced) atomicAdd(P.counters + 0, cubes_traced);
}

}  // namespace aicb
#endif  // __CUDACC__
