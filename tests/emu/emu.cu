// TEST INFRASTRUCTURE ONLY — host-side emulator of the device math.
//
// Compiles legged-robot-movability-cuda_b200/csrc/leg_math.cuh (the exact source the CUDA kernels
// are built from) for the HOST, so that the CPU test suite can compare the product's per-point
// arithmetic and its leg-plan construction against the oracle without a GPU.  It is not part of
// liblrm_b200.so, is not reachable from the C ABI, and is never used by bench.py: it exists to
// catch algorithmic regressions before GPU minutes are spent.  Differences from the device:
// rsqrtf is 1/sqrtf here (a few ulp), everything else is the same FP32 expression tree.
#include <cstdint>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "leg_math.cuh"
#include "leg_plan.h"

namespace {
void host_table(const lrm::LegPlan& L, lrm::SectorTable* tab) {
    lrm::fill_sector_table(L, tab, 0, 1);
}
}  // namespace

extern "C" {

void emu_reach(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, uint8_t* out) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        out[i] = lrm::reach_coxa_frame(L, tab, p) ? 1 : 0;
    }
}

void emu_dist(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, float* out_vec,
              uint8_t* out_flag, uint8_t* out_reach) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        const lrm::DistResult r = L.generic ? lrm::dist_coxa_frame<true>(L, tab, p) : lrm::dist_coxa_frame<false>(L, tab, p);
        out_vec[3 * i] = r.dx, out_vec[3 * i + 1] = r.dy, out_vec[3 * i + 2] = r.dz;
        if (out_flag) out_flag[i] = r.flag ? 1 : 0;
        if (out_reach) out_reach[i] = r.reach ? 1 : 0;
    }
}

// the reachable_rotate_leg predicate as the positionability kernel evaluates it
void emu_leg_reaches(const float* offsets, size_t n, const lrm_leg_t* leg, const float* quat,
                     uint8_t* out) {
    lrm::LegPlan L;
    lrm::build_leg_plan_rotated_limits(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    for (size_t i = 0; i < n; i++) {
        const float vx = offsets[3 * i], vy = offsets[3 * i + 1], vz = offsets[3 * i + 2];
        const float g = fmaf(L.grav[0], vx, fmaf(L.grav[1], vy, L.grav[2] * vz));
        out[i] = (!(g < 0.f) && lrm::reach_coxa_frame(L, tab, lrm::to_coxa_frame(L, vx, vy, vz))) ? 1 : 0;
    }
}

// same points through the explicit cross-validation path (the fallback for exotic legs)
void emu_dist_generic(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat,
                      float* out_vec) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        const lrm::DistResult r = lrm::dist_coxa_frame<true>(L, tab, p);
        out_vec[3 * i] = r.dx, out_vec[3 * i + 1] = r.dy, out_vec[3 * i + 2] = r.dz;
    }
}

// The fast path exactly as the streaming kernel runs it (dist_fast: yaw-sector table + plane
// atlas); points the tables cannot certify fall back to the full evaluation.  Returns the number
// of points that fell back; *pure_cells / *pure_bins report the certified share of both tables.
size_t emu_dist_atlas(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, int dim,
                      float cell, float* out_vec, uint8_t* out_flag, uint8_t* out_reach,
                      size_t* pure_cells, size_t* pure_bins, uint8_t* out_reach_atlas) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    lrm::WinnerTable win;
    lrm::fill_winner_table(L, &win, 0, 1);
    lrm::FastTables ft;
    lrm::build_fast_tables(L, &ft);
    if (pure_bins) {
        *pure_bins = 0;
        for (int b = 0; b <= lrm::kYawBins; b++) *pure_bins += ft.code[b] != lrm::kYawImpure;
    }
    const float origin = -0.5f * dim * cell;
    unsigned char* cells = new unsigned char[(size_t)dim * dim];
    const float need = lrm::kAtlasNeedFactor * cell + lrm::kAtlasNeedSlack;  // rule of atlas_build_kernel
    size_t pure = 0;
    for (int iy = 0; iy < dim; iy++)
        for (int ix = 0; ix < dim; ix++) {
            const float X = origin + ((float)ix + 0.5f) * cell, Y = origin + ((float)iy + 0.5f) * cell;
            const lrm::PlaneProbe pr = lrm::plane_probe(L, tab, X, Y);
            const bool ok = pr.safety > need;
            cells[lrm::atlas_index(dim, ix, iy)] = (unsigned char)lrm::atlas_cell_byte(pr, need);
            pure += ok;
        }
    if (pure_cells) *pure_cells = pure;
    lrm::AtlasView A{cells, 0, 1.0f / cell, -origin / cell, -origin / cell, dim, dim};
    const lrm::FastView F{ft.pair, ft.code, ft.combo, ft.ncombo};
    size_t fallback = 0;
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        if (out_reach_atlas) out_reach_atlas[i] = lrm::reach_coxa_frame_atlas<false>(L, tab, A, p) ? 1 : 0;
        lrm::DistResult r;
        if (!lrm::dist_fast<false>(L, F, A, win, p, &r)) {
            r = lrm::dist_coxa_frame<false>(L, tab, p);
            fallback++;
        }
        out_vec[3 * i] = r.dx, out_vec[3 * i + 1] = r.dy, out_vec[3 * i + 2] = r.dz;
        if (out_flag) out_flag[i] = r.flag ? 1 : 0;
        if (out_reach) out_reach[i] = r.reach ? 1 : 0;
    }
    delete[] cells;
    return fallback;
}

// The three-tier distance sweep as the streaming kernel runs it: choice volume (dist_choice) ->
// dist_fast -> full evaluation.  Cube bytes are computed on demand with the very function the
// device build kernel uses (choice_cell_byte).  tiers[0..3] = points decided by tier 1, dist_fast, dist_choice_clamp, the full evaluation.
void emu_dist_choice(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, int dim,
                     float cell, float vol_h, int vol_dim, float* out_vec, uint8_t* out_flag,
                     uint8_t* out_reach, size_t* tiers, uint8_t* out_tier, uint8_t* out_reach_vol,
                     size_t* reach_known) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    lrm::WinnerTable win;
    lrm::fill_winner_table(L, &win, 0, 1);
    lrm::FastTables ft;
    lrm::build_fast_tables(L, &ft);
    if (reach_known) *reach_known = 0;
    const float origin = -0.5f * dim * cell;
    std::vector<unsigned char> cells((size_t)dim * dim);
    const float need = lrm::kAtlasNeedFactor * cell + lrm::kAtlasNeedSlack;
    for (int iy = 0; iy < dim; iy++)
        for (int ix = 0; ix < dim; ix++) {
            const float X = origin + ((float)ix + 0.5f) * cell, Y = origin + ((float)iy + 0.5f) * cell;
            cells[lrm::atlas_index(dim, ix, iy)] = (unsigned char)lrm::atlas_cell_byte(lrm::plane_probe(L, tab, X, Y), need);
        }
    lrm::AtlasView A{cells.data(), 0, 1.0f / cell, -origin / cell, -origin / cell, dim, dim};
    const lrm::FastView F{ft.pair, ft.code, ft.combo, ft.ncombo};
    const lrm::YawSol* sols = reinterpret_cast<const lrm::YawSol*>(ft.pair);
    std::unordered_map<uint64_t, unsigned char> cubes;
    const float vo = 0.5f * vol_dim, vinv = 1.0f / vol_h;
    tiers[0] = tiers[1] = tiers[2] = tiers[3] = 0;
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        const float fx = fmaf(p.x, vinv, vo), fy = fmaf(p.y, vinv, vo + lrm::kVolShiftY), fz = fmaf(p.z, vinv, vo);
        unsigned cube = 0;
        if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)vol_dim && fy < (float)vol_dim && fz < (float)vol_dim) {
            const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
            const uint64_t key = ((uint64_t)iz * vol_dim + iy) * vol_dim + ix;
            auto it = cubes.find(key);
            if (it == cubes.end()) {
                const unsigned char b = (unsigned char)lrm::choice_cell_byte(
                    L, tab, ft, ((float)ix - vo) * vol_h, ((float)iy - vo - lrm::kVolShiftY) * vol_h, ((float)iz - vo) * vol_h, vol_h);
                it = cubes.emplace(key, b).first;
            }
            cube = it->second;
        }
        if (out_reach_vol) {  // the reach-only sweep: reach bits of the cube, else the atlas path
            if (cube & lrm::kVolReachKnown) {
                out_reach_vol[i] = (cube & lrm::kVolReachValue) ? 1 : 0;
                if (reach_known) ++*reach_known;
            } else {
                out_reach_vol[i] = lrm::reach_coxa_frame_atlas<false>(L, tab, A, p) ? 1 : 0;
            }
        }
        lrm::DistResult r;
        int tier = 0;
        const int st = lrm::dist_choice<false>(L, sols, cube, A, win, p, &r);
        if (st == 2) {  // cube certified, plane cell not: the chosen solution, explicit plane evaluation
            lrm::dist_choice_clamp(L, tab, sols, cube, p, &r);
            tier = 2;
        } else if (st == 1) {
            tier = 1;
            if (!lrm::dist_fast<false, false, true>(L, F, A, win, p, &r)) {  // as ring B's redo runs it
                r = lrm::dist_coxa_frame<false>(L, tab, p);
                tier = 3;
            }
        }
        tiers[tier]++;
        if (out_tier) out_tier[i] = (uint8_t)tier;
        out_vec[3 * i] = r.dx, out_vec[3 * i + 1] = r.dy, out_vec[3 * i + 2] = r.dz;
        if (out_flag) out_flag[i] = r.flag ? 1 : 0;
        if (out_reach) out_reach[i] = r.reach ? 1 : 0;
    }
}

// The tiered distance sweep WITH tier 0 (16-bit cube texels: chosen solution + plane label), as
// one_leg_tier_kernel runs it: dist_choice_label -> dist_choice (plane atlas) -> dist_choice_clamp /
// dist_fast -> full evaluation.  Texels are computed on demand the way the device build does it:
// the block of 4^3 cubes first (probe at the block centre), the cube itself (atlas scan) if the
// block is not settled as a whole.  tiers[0..4] = points decided by tier 0, the atlas tier, the
// explicit plane evaluation, dist_fast, the full evaluation.
void emu_dist_tier0(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, int dim, float cell,
                    float vol_h, int vol_dim, float* out_vec, uint8_t* out_flag, uint8_t* out_reach,
                    size_t* tiers, uint8_t* out_tier, int bricks) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    lrm::WinnerTable win;
    lrm::fill_winner_table(L, &win, 0, 1);
    lrm::FastTables ft;
    lrm::build_fast_tables(L, &ft);
    const float origin = -0.5f * dim * cell;
    std::vector<unsigned char> cells((size_t)dim * dim);
    const float need = lrm::kAtlasNeedFactor * cell + lrm::kAtlasNeedSlack;
    for (int iy = 0; iy < dim; iy++)
        for (int ix = 0; ix < dim; ix++) {
            const float X = origin + ((float)ix + 0.5f) * cell, Y = origin + ((float)iy + 0.5f) * cell;
            cells[lrm::atlas_index(dim, ix, iy)] = (unsigned char)lrm::atlas_cell_byte(lrm::plane_probe(L, tab, X, Y), need);
        }
    lrm::AtlasView A{cells.data(), 0, 1.0f / cell, -origin / cell, -origin / cell, dim, dim};
    const lrm::FastView F{ft.pair, ft.code, ft.combo, ft.ncombo};
    const lrm::YawSol* sols = reinterpret_cast<const lrm::YawSol*>(ft.pair);
    std::unordered_map<uint64_t, unsigned> cubes, blocks;
    const float vo = 0.5f * vol_dim, vinv = 1.0f / vol_h;
    for (int k = 0; k < 5; k++) tiers[k] = 0;
    std::unordered_map<uint64_t, unsigned> fine;
    // classic texel of a cube, computed on demand the way the device build does it
    auto classic_word = [&](int ix, int iy, int iz) -> unsigned {
        const uint64_t bkey = ((uint64_t)(iz >> 2) * vol_dim + (iy >> 2)) * vol_dim + (ix >> 2);
        auto bt = blocks.find(bkey);
        if (bt == blocks.end()) {
            // volume_coarse_kernel: the block is settled only if nothing needs refinement
            const float x0 = ((float)(ix & ~3) - vo) * vol_h, y0 = ((float)(iy & ~3) - vo - lrm::kVolShiftY) * vol_h,
                        z0 = ((float)(iz & ~3) - vo) * vol_h;
            bt = blocks.emplace(bkey, lrm::coarse_block_word(L, tab, ft, x0, y0, z0, 4.f * vol_h)).first;
        }
        if (bt->second != 0u) return bt->second;
        const uint64_t key = ((uint64_t)iz * vol_dim + iy) * vol_dim + ix;
        auto it = cubes.find(key);
        if (it == cubes.end())
            it = cubes.emplace(key, lrm::choice_cell_word(L, tab, ft, A, ((float)ix - vo) * vol_h,
                                                          ((float)iy - vo - lrm::kVolShiftY) * vol_h,
                                                          ((float)iz - vo) * vol_h, vol_h, false)).first;
        return it->second;
    };
    for (size_t i = 0; i < n; i++) {
        const lrm::CoxaPoint p = lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        const float fx = fmaf(p.x, vinv, vo), fy = fmaf(p.y, vinv, vo + lrm::kVolShiftY), fz = fmaf(p.z, vinv, vo);
        unsigned word = 0;
        const bool inside = fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)vol_dim && fy < (float)vol_dim && fz < (float)vol_dim;
        if (inside) word = classic_word((int)fx, (int)fy, (int)fz);
        if (bricks && inside) {
            // bricks (leg_math.cuh): an unsettled cube points to 4^3 fine cubes, certified by the same
            // functions over boxes widened by the COARSE pad.  The device finds the brick through a
            // texture fetch that resolves cube coordinates to 1 / 256 of a cube: next to a face it may
            // pick the neighbouring cube.  Both outcomes are emulated; where both yield a tier-0 texel
            // the results must agree bit for bit.
            const lrm::VolumeView V{0, vinv, vo, vo + lrm::kVolShiftY, vol_dim, nullptr};
            auto texel_via = [&](int cx, int cy, int cz) -> unsigned {
                if (cx < 0 || cy < 0 || cz < 0 || cx >= vol_dim || cy >= vol_dim || cz >= vol_dim) return 0u;
                const unsigned cw = classic_word(cx, cy, cz);
                if (!lrm::brick_candidate(cw)) return cw;
                const unsigned ptr = lrm::brick_texel(0u, cx, cy, cz, cw);
                const unsigned slot = lrm::brick_slot(V, ptr, p);
                const uint64_t key = ((((uint64_t)cz * vol_dim + cy) * vol_dim + cx) << 6) | slot;
                auto it = fine.find(key);
                if (it == fine.end())
                    it = fine.emplace(key, lrm::brick_fine_word(L, tab, ft, A, cw, ((float)cx - vo) * vol_h,
                                                                ((float)cy - vo - lrm::kVolShiftY) * vol_h,
                                                                ((float)cz - vo) * vol_h, vol_h, slot)).first;
                return it->second;
            };
            const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
            word = texel_via(ix, iy, iz);
            const float q = 1.0f / 200.f;  // a little more than the texture unit's 1 / 256
            for (int ax = 0; ax < 3; ax++) {
                const float f = ax == 0 ? fx : ax == 1 ? fy : fz;
                const float fr = f - floorf(f);
                const int d = fr < q ? -1 : fr > 1.f - q ? 1 : 0;
                if (d == 0) continue;
                const unsigned other = texel_via(ix + (ax == 0 ? d : 0), iy + (ax == 1 ? d : 0), iz + (ax == 2 ? d : 0));
                lrm::DistResult ra, rb;
                if (lrm::dist_choice_label<true>(L, sols, word, win, p, &ra) == 0 &&
                    lrm::dist_choice_label<true>(L, sols, other, win, p, &rb) == 0 &&
                    (std::memcmp(&ra.dx, &rb.dx, 12) != 0 || ra.flag != rb.flag || ra.reach != rb.reach))
                    tiers[0] = (size_t)-1 << 20;
                if (lrm::dist_choice_label<true>(L, sols, word, win, p, &ra) != 0) word = other;  // exercise the neighbour's texel too
            }
        }
        lrm::DistResult r;
        int tier = 0;
        int st = lrm::dist_choice_label<true>(L, sols, word, win, p, &r);
        if (st == 0 && !(word & 0x4000u)) {
            // a warp without any valid plane point takes the version without the limit-plane rule
            lrm::DistResult r2;
            lrm::dist_choice_label<false>(L, sols, word, win, p, &r2);
            if (std::memcmp(&r2.dx, &r.dx, 12) != 0 || r2.flag != r.flag || r2.reach != r.reach) tiers[0] = (size_t)-1 << 20;
        }
        if (st == 3) {  // ring A0: the plane atlas; its cell may be uncertified too (ring A)
            tier = 1;
            st = lrm::dist_choice<false>(L, sols, (word & 0xffu) | lrm::kVolPure, A, win, p, &r);
            if (st == 2) {
                lrm::dist_choice_clamp(L, tab, sols, word & 0xffu, p, &r);
                tier = 2;
            }
        } else if (st == 1) {
            tier = 3;
            if (!lrm::dist_fast<false, false, true>(L, F, A, win, p, &r)) {
                r = lrm::dist_coxa_frame<false>(L, tab, p);
                tier = 4;
            }
        }
        tiers[tier]++;
        if (out_tier) out_tier[i] = (uint8_t)tier;
        out_vec[3 * i] = r.dx, out_vec[3 * i + 1] = r.dy, out_vec[3 * i + 2] = r.dz;
        if (out_flag) out_flag[i] = r.flag ? 1 : 0;
        if (out_reach) out_reach[i] = r.reach ? 1 : 0;
    }
}

// The positionability kernel's per-pose logic (positionability.cu) with brute-force target loops:
// same orientation matrix, same cull cylinders, same leg predicate arithmetic.
void emu_standability(const float* bodies, size_t nb, const float* map, size_t nt,
                      const lrm_leg_t* legs, int nlegs, const float* quats, int nq, uint8_t* out) {
    const float pi = 3.14159265358979323846264338327950288419716939937510582097f;
    std::vector<lrm::ReachPlan> plans((size_t)nq * nlegs);
    struct OC { float R[9], radius_in, plus_in, minus_in, radius_out; lrm::GravExact grav; };
    std::vector<OC> oc(nq);
    for (int o = 0; o < nq; o++) {
        const float* q = quats + 4 * o;
        const float ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
        float c0[3], c1[3], c2[3];
        lrm::quat_rotate(q, ex, c0), lrm::quat_rotate(q, ey, c1), lrm::quat_rotate(q, ez, c2);
        for (int r = 0; r < 3; r++) oc[o].R[3 * r] = c0[r], oc[o].R[3 * r + 1] = c1[r], oc[o].R[3 * r + 2] = c2[r];
        for (int l = 0; l < nlegs; l++) {
            lrm::LegPlan full;
            lrm::build_leg_plan_rotated_limits(legs[l], q, &full);
            float az_s, az_c;
            sincosf(-legs[l].body_angle, &az_s, &az_c);
            lrm::make_reach_plan(full, legs[l].min_angle_coxa, legs[l].max_angle_coxa, az_c, az_s,
                                 &plans[(size_t)o * nlegs + l]);
        }
        lrm::make_grav_exact(q, &oc[o].grav);
        lrm_leg_t d = legs[0];
        const float pitch = lrm::quat_pitch_for_leg(q, d.body_angle);
        d.tibia_absolute_pos -= pitch, d.tibia_absolute_neg -= pitch;
        const float s_p = std::sin(d.coxa_pitch), c_p = std::cos(d.coxa_pitch);
        oc[o].radius_in = d.body + c_p * d.coxa_length + d.femur_length + d.tibia_length;
        const float plus_abs = d.tibia_length * std::sin(d.tibia_absolute_pos) +
                               d.femur_length * std::sin(std::min(pi / 2, d.max_angle_femur));
        oc[o].plus_in = s_p * d.coxa_length + plus_abs;
        oc[o].minus_in = s_p * d.coxa_length - d.femur_length - d.tibia_length;
        oc[o].radius_out = d.body;
    }
    auto rot = [](const float* R, float x, float y, float z, float* o) {
        o[0] = fmaf(R[0], x, fmaf(R[1], y, R[2] * z));
        o[1] = fmaf(R[3], x, fmaf(R[4], y, R[5] * z));
        o[2] = fmaf(R[6], x, fmaf(R[7], y, R[8] * z));
    };
    std::vector<float> T(3 * nt);
    std::vector<uint8_t> res(nb, 0);
    for (int o = 0; o < nq; o++) {
        for (size_t t = 0; t < nt; t++) rot(oc[o].R, map[3 * t], map[3 * t + 1], map[3 * t + 2], &T[3 * t]);
        for (size_t b = 0; b < nb; b++) {
            if (res[b]) continue;
            float B[3];
            rot(oc[o].R, bodies[3 * b], bodies[3 * b + 1], bodies[3 * b + 2], B);
            bool near = false, hit = false;
            for (size_t t = 0; t < nt && !hit; t++) {
                const float dz = T[3 * t + 2] - B[2];
                const float dx = T[3 * t] - B[0], dy = T[3 * t + 1] - B[1];
                const float rad = sqrtf(dx * dx + dy * dy);
                if (rad < oc[o].radius_in && dz < oc[o].plus_in && dz > oc[o].minus_in) near = true;
                if (rad < oc[o].radius_out && dz < 250.f && dz > -110.f) hit = true;
            }
            if (hit || !near) continue;
            bool all = true;
            for (int l = 0; l < nlegs && all; l++) {
                const lrm::ReachPlan& L = plans[(size_t)o * nlegs + l];
                bool found = false;
                for (size_t t = 0; t < nt && !found; t++) {
                    const float vx = T[3 * t] - B[0], vy = T[3 * t + 1] - B[1], vz = T[3 * t + 2] - B[2];
                    if (fabsf(vx) > 520.f || fabsf(vy) > 520.f) continue;
                    const lrm::GravCtx gc{&oc[o].grav, bodies[3 * b], bodies[3 * b + 1], bodies[3 * b + 2],
                                          map[3 * t], map[3 * t + 1], map[3 * t + 2]};
                    if (lrm::reach_offset(L, vx, vy, vz, &gc)) found = true;
                }
                all = found;
            }
            if (all) res[b] = (uint8_t)(o + 1);
        }
    }
    std::memcpy(out, res.data(), nb);
}

// ReachPlan predicates: out_reach[i] = reach_offset(v_i); out_ball[i] = reach_ball_possible(v_i, rc)
void emu_reach_offset(const float* offsets, size_t n, const lrm_leg_t* leg, const float* quat, float rc,
                      uint8_t* out_reach, uint8_t* out_ball) {
    lrm::LegPlan full;
    lrm::build_leg_plan_rotated_limits(*leg, quat, &full);
    lrm::ReachPlan L;
    float az_s, az_c;
    sincosf(-leg->body_angle, &az_s, &az_c);
    lrm::make_reach_plan(full, leg->min_angle_coxa, leg->max_angle_coxa, az_c, az_s, &L);
    for (size_t i = 0; i < n; i++) {
        const float vx = offsets[3 * i], vy = offsets[3 * i + 1], vz = offsets[3 * i + 2];
        out_reach[i] = lrm::reach_offset(L, vx, vy, vz) ? 1 : 0;
        out_ball[i] = lrm::reach_ball_possible(L, vx, vy, vz, rc) ? 1 : 0;
    }
}

// leg_ball_possible (the octree's per-leg pruning test) for world points (full plan with `quat`)
void emu_leg_ball(const float* xyz, size_t n, const lrm_leg_t* leg, const float* quat, float rc, uint8_t* out) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    const bool wedge = leg->max_angle_coxa >= leg->min_angle_coxa && leg->max_angle_coxa - leg->min_angle_coxa < 3.0f;
    for (size_t i = 0; i < n; i++)
        out[i] = lrm::leg_ball_possible(L, lrm::to_coxa_frame(L, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), rc, wedge) ? 1 : 0;
}

// The certificate of the plane atlas, checked directly: plane_probe(P).safety is a radius within which
// the outcome of the plane evaluation — valid bit, sector, winning candidate — cannot change.  For each
// of the n centres (X, Y) with a positive safety, `samples` points at 0 .. 0.98 x safety (capped at
// `cap` mm) in pseudo-random directions must evaluate to the same label, and plane_clamp itself must
// pick the labelled winner there (its result equals plane_from_label's, bit for bit).  Returns the
// number of violations; counts[0] = centres with safety > 0, counts[1] = samples checked,
// counts[2] = centres certified only thanks to the corner-on-circle certificate.
size_t emu_probe_ball_check(const float* centres, size_t n, const lrm_leg_t* leg, const float* quat, int samples,
                            float cap, size_t* counts) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    lrm::SectorTable tab;
    host_table(L, &tab);
    lrm::WinnerTable win;
    lrm::fill_winner_table(L, &win, 0, 1);
    size_t bad = 0;
    counts[0] = counts[1] = counts[2] = 0;
    uint32_t rng = 12345u;
    auto next = [&]() { rng = rng * 1664525u + 1013904223u; return (float)(rng >> 8) * (1.0f / 16777216.0f); };
    for (size_t i = 0; i < n; i++) {
        const float X = centres[2 * i], Y = centres[2 * i + 1];
        const lrm::PlaneProbe pr = lrm::plane_probe(L, tab, X, Y);
        if (!(pr.safety > 0.05f) || (pr.label & 15) == lrm::kAtlasNone) continue;
        counts[0]++;
        const float rmax = fminf(pr.safety, cap) * 0.98f;
        for (int k = 0; k < samples; k++) {
            const float ang = 6.2831853f * next(), rad = rmax * sqrtf(next());
            const float x = X + rad * cosf(ang), y = Y + rad * sinf(ang);
            counts[1]++;
            const lrm::PlaneProbe q = lrm::plane_probe(L, tab, x, y);
            const lrm::PlaneResult a = lrm::plane_clamp<false>(L, tab, x, y);
            const lrm::PlaneResult b = lrm::plane_from_label(win, (unsigned)pr.label, x, y);
            if (q.label != pr.label || a.valid != b.valid || std::memcmp(&a.dx, &b.dx, 4) != 0 ||
                std::memcmp(&a.dy, &b.dy, 4) != 0)
                bad++;
        }
    }
    return bad;
}

// The pose search's orientation-independent collision region (leg_math.cuh: AxisCone): claimed[i] = the
// region contains offset i; truth[i] = the offset is inside the body cylinder (radius `radius`, heights
// in (-110, 250), several_leg.cu:504-559) under EVERY one of the nq orientations, evaluated the way
// positionability.cu does it (matrix from the rotated basis vectors, T = R d).  Returns cone.ok | gate << 1.
int emu_cone_check(const float* quats, int nq, float radius, const float* offsets, size_t n, uint8_t* claimed,
                   uint8_t* truth) {
    std::vector<float> R((size_t)nq * 9), axes((size_t)nq * 3);
    const float ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1};
    for (int o = 0; o < nq; o++) {
        float c0[3], c1[3], c2[3];
        lrm::quat_rotate(quats + 4 * o, ex, c0), lrm::quat_rotate(quats + 4 * o, ey, c1), lrm::quat_rotate(quats + 4 * o, ez, c2);
        for (int r = 0; r < 3; r++) R[9 * o + 3 * r] = c0[r], R[9 * o + 3 * r + 1] = c1[r], R[9 * o + 3 * r + 2] = c2[r];
        axes[3 * o] = R[9 * o + 6], axes[3 * o + 1] = R[9 * o + 7], axes[3 * o + 2] = R[9 * o + 8];
    }
    lrm::AxisCone C;
    lrm::make_axis_cone(axes.data(), nq, -110.f, 250.f, radius, &C);
    for (size_t i = 0; i < n; i++) {
        const float dx = offsets[3 * i], dy = offsets[3 * i + 1], dz = offsets[3 * i + 2];
        claimed[i] = (C.ok && lrm::cone_collides_always(C, dx, dy, dz)) ? 1 : 0;
        bool all = true;
        for (int o = 0; o < nq && all; o++) {
            const float* M = &R[9 * o];
            const float tx = fmaf(M[0], dx, fmaf(M[1], dy, M[2] * dz)), ty = fmaf(M[3], dx, fmaf(M[4], dy, M[5] * dz)),
                        tz = fmaf(M[6], dx, fmaf(M[7], dy, M[8] * dz));
            all = fmaf(tx, tx, ty * ty) < radius * radius && tz < 250.f && tz > -110.f;
        }
        truth[i] = all ? 1 : 0;
    }
    return C.ok | (C.gate << 1);
}

int emu_plan_is_generic(const lrm_leg_t* leg, const float* quat) {
    lrm::LegPlan L;
    lrm::build_leg_plan(*leg, quat, &L);
    return L.generic;
}

int emu_sizeof_plan() { return (int)sizeof(lrm::LegPlan); }

// the host-built plan and yaw tables, byte for byte (pinned by tests/golden/leg_plan_golden.npz)
void emu_plan_bytes(const lrm_leg_t* leg, const float* quat, unsigned char* plan_out, unsigned char* tables_out) {
    lrm::LegPlan L;
    std::memset(&L, 0, sizeof L);
    lrm::build_leg_plan(*leg, quat, &L);
    std::memcpy(plan_out, &L, sizeof L);
    lrm::FastTables ft;
    std::memset(&ft, 0, sizeof ft);
    lrm::build_fast_tables(L, &ft);
    std::memcpy(tables_out, &ft, sizeof ft);
}
int emu_sizeof_tables() { return (int)sizeof(lrm::FastTables); }

}  // extern "C"
