// leg_plan.h — everything about one (leg, body orientation) pair that does not depend on the
// query point, computed once on the host and handed to the kernels as a __grid_constant__
// parameter (constant bank: operands feed FFMA directly, no registers, no per-point trig).
//
// The reference recomputes all of this per point: 6 circles + up to 10 corner points, each with
// sin/cos, into a Circle[14] local-memory array (one_leg.cu:175-180, circles.cu.h:80-135,337-383,
// 417-476), plus rotate_leg_data per block/point (one_leg_global.cu:48-60,80-87).  LegCompact
// (HeaderCPP.h:54-76) was the reference's unfinished attempt at the same idea.
#pragma once
#include <stdint.h>

#include "lrm_c.h"

namespace lrm {

// Decides  atan2f(Y, X) > theta  without evaluating atan2f:
//   cr = c*Y + ns*X + bias      (c = cos theta, ns = -sin theta)
//   up = !signbit(Y)            (angle in [0, pi])
//   theta in [0, pi):  up && cr > 0        theta in [-pi, 0):  up || cr > 0
// bias is 0 except for thresholds outside (-pi, pi): always-true / always-false.
// "angle < theta" is built as  atan2f(-Y, X) > -theta  (atan2f is odd in Y, signed zeros included).
// Branch-free device form:  cr > thr  with  thr = thr_dn - up * kAngleBig, where
//   thr_dn = kAngleBig for the first kind (so thr = 0 when up, +big otherwise) and
//   thr_dn = 0         for the second     (so thr = -big when up, 0 otherwise).
constexpr float kAngleBig = 1.0e30f;
struct AngleTest {
    float c, ns, bias;
    uint32_t lower;
    float thr_dn;
};

// Circle in the femur plane with its validity rule folded into one compare:
//   valid  <=>  sgn * |P - centre|^2  <  thr_s
// attractive (must be inside, +-CIRCLE_MARGIN slack, one_leg.cu:31-41): sgn = +1, thr_s = (r+eps)^2
// repulsive  (must be outside):                                        sgn = -1, thr_s = -(r-eps)^2
struct PlanCircle {
    float cx, cy, r, sgn, thr_s;
};

constexpr int kMaxCorners = 10;

// One row of the 4-sector table the kernels stage in shared memory (sector = upper*2 + ext).
// Per slot j = 1..3 eight floats: cx, cy, r, sgn, ax, ay, ah, spare.
//   (ax, ay, ah): the part of circle j that satisfies the OTHER circles of the sector (the
//   cross-validation of multi_circle_clamp, one_leg.cu:122-123) is an arc, stored as its bisector
//   direction and the cosine of its half-angle:  projection valid  <=>  (P-c).(ax,ay) >= ah*|P-c|
//   (ah = 2: never, ah = -2: always).
//   spare of slots 1,2,3: the same three numbers for the inner circle (slot 0) of this sector,
//   whose valid set can consist of two arcs; inner_b holds the second one (empty: ah = 2).
struct alignas(16) SectorRow {
    float slot[3][8];
    float inner_b[4];
};

struct LegPlan {
    // world point -> coxa frame (qtInvRotate, Rz(-body_angle), x -= body, Ry(-coxa_pitch):
    // one_leg_global.cu:88-95,119-127, one_leg.cu:9-24) as one affine map p' = M p + t
    float M[9];
    float t[3];
    // coxa-frame vector -> world (Ry(+pitch), Rz(+body_angle), qtRotate: one_leg.cu:339,
    // one_leg_global.cu:97-99)
    float Mo[9];
    // gravity-side half-space of reachable_rotate_leg (several_leg.cu:58-62): a foothold offset v
    // (orientation frame) is rejected when grav . v < 0
    float grav[3];
    float coxa_length;

    // coxa yaw limits (one_leg.cu:222-234,305-306)
    AngleTest over;     // a > max_angle_coxa
    AngleTest under;    // a < min_angle_coxa        (apply to (X, -Y))
    AngleTest mega_hi;  // a > max_angle_coxa + pi/2
    AngleTest mega_lo;  // a < min_angle_coxa - pi/2 (apply to (X, -Y))
    AngleTest mid;      // a > (max + min) / 2
    float cos_max, sin_max, cos_min, sin_min;

    // femur-plane sectors (find_region, circles.cu.h:48-78)
    AngleTest middle;   // angle > middle_angle  -> UpperRegion
    AngleTest sat[2];   // [UpperRegion]: angle > saturation limit of that side

    // circle sets (insert_circles, circles.cu.h:337-383): slot 0 is the inner circle for every
    // sector; slots 1..3 depend on UpperRegion only, except that the attractive one becomes the
    // outer circle when FullyExtended.
    PlanCircle inner;
    PlanCircle outer;          // attractive form of the outer circle
    PlanCircle slot[2][3];     // [UpperRegion][slot-1]
    int32_t att_slot[2];       // which of slot[u][0..2] is the attractive one

    // host-built shared-memory table (see SectorRow) + whether every valid set is a single arc;
    // if not (exotic leg), the kernels fall back to explicit cross-validation (generic != 0).
    SectorRow sector[4];
    int32_t generic;
    int32_t std_coxa;  // all five coxa yaw tests are the plain "upper, no bias" form

    // corner points of the planar workspace (insert_intersecv2, circles.cu.h:417-476), in
    // emission order (ties keep the earlier candidate, one_leg.cu:133)
    int32_t n_corners;
    float corner_x[kMaxCorners];
    float corner_y[kMaxCorners];
};

// ---- yaw sectors (fast path of the distance sweep) -------------------------------------------
// Every yaw decision of finish_finding_closest (one_leg.cu:222-234: beyond limit +- pi/2, over /
// under the limits, which limit plane the "coxa-limit plane is nearer" rule uses) depends on the
// direction of the point around the coxa axis only.  The circle of directions is cut into
// kYawBins bins of equal "diamond angle" (uy / (|ux| + |uy|), monotone in atan2); a bin whose five
// yaw tests are constant for the direct AND the pi-flipped solution over the whole (padded) bin
// carries the index of its (direct, flipped) solution pair in `pair`; bins that contain a
// decision boundary, the +-pi seam or the x axis (signed-zero rules) are kYawImpure and the point
// takes the full evaluation.  The table never changes a result, it only replaces ~100
// instructions of tests and selects by one shared-memory load.
constexpr int kYawBins = 1024;
constexpr int kYawPairs = 16;         // distinct (direct, flipped) combinations a leg can have
constexpr uint8_t kYawImpure = 0xFF;
struct alignas(16) YawSol {
    float k;           // plane direction (cs, ss) = k * (ux, uy) + (c_cs, c_ss):
    float c_cs, c_ss;  //   k = +-1 for an unsaturated / mega-saturated yaw, 0 at a coxa limit
    float nsat;        // 1 if the yaw is unsaturated (res = valid), else 0
    float cl, sl;      // direction of the coxa limit used by one_leg.cu:258-274
    float big;         // 0, or +inf when that rule is off (mega-saturated)
    float present;     // 0 for a solution that duplicates the other one (one_leg.cu:225-226)
};
struct YawPair {
    YawSol a, b;  // direct and pi-flipped solution of one yaw sector
};
struct FastTables {
    YawPair pair[kYawPairs];
    uint8_t code[kYawBins + 16];  // bins 0 .. kYawBins: index into pair, or kYawImpure
    // the (direct kind, flipped kind) combination each pair stands for (see yaw_combo), so that
    // the choice volume can map a direction's yaw decisions to its pair without going through bins
    int32_t ncombo;
    int32_t combo[kYawPairs];
    int32_t both_unsat;  // some pair has two unsaturated solutions (yaw range >= pi): no choice volume
};
void build_fast_tables(const LegPlan& plan, FastTables* out);

// Host-side construction.  quat may be nullptr (identity).  Pure FP arithmetic, no CUDA.
void build_leg_plan(const lrm_leg_t& leg, const float* quat, LegPlan* out);
// Variant used by the positionability path (several_leg.cu:48-67,743-760): the tibia limits are
// shifted by the pitch of `quat` seen from the leg azimuth, but points arrive already expressed
// in the orientation frame, so the point transform is only Rz(-body_angle) etc.
void build_leg_plan_rotated_limits(const lrm_leg_t& leg, const float* quat, LegPlan* out);

// The reference's quaternion helpers (unified_math_cuda.cu.h:13-83, octree_util.cu.h:164-172) in
// their original mixed storage layouts; host only.
void quat_from_vect_angle(const float axis[3], float angle, float out[4]);
void quat_multiply(const float a[4], const float b[4], float out[4]);
void quat_invert(const float q[4], float out[4]);
void quat_rotate(const float q[4], const float v[3], float out[3]);
float quat_pitch_for_leg(const float quat[4], float body_angle);  // rotate_leg_data's pitch
void default_leg(int robot, float azimuth, lrm_leg_t* out);       // static_variables.cpp:6-93

}  // namespace lrm
