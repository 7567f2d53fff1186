// leg_plan.cpp — host-side construction of LegPlan (see leg_plan.h).
//
// Everything here runs once per (leg, orientation), so clarity wins over speed.  Leg constants
// that the reference evaluates in float (circle centres, radii, the oriented tibia limits) are
// evaluated in float with the same formulas so they agree to the last bit with what the
// reference's host path uses; derived thresholds (squared radii with margin, directions of the
// angle thresholds, the fused affine maps) are evaluated in double and rounded once.
#include "leg_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace lrm {

namespace {
constexpr float kPi = 3.14159265358979323846264338327950288419716939937510582097f;
constexpr double kCircleMargin = 0.001;  // settings.h:9 (a double literal in the reference)
constexpr double kCornerEps = 0.001;     // circles.cu.h:7

struct Mat3 {
    double m[9];
};
Mat3 mul(const Mat3& a, const Mat3& b) {
    Mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += a.m[3 * i + k] * b.m[3 * k + j];
            r.m[3 * i + j] = s;
        }
    return r;
}
// Matrix of v -> qtRotate(q, v) (unified_math_cuda.cu.h:13-27); linear in v for any q.
Mat3 quat_matrix(const float q[4]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double t2 = x * y, t3 = x * z, t4 = x * w, t5 = -y * y, t6 = y * z, t7 = y * w,
                 t8 = -z * z, t9 = z * w, t10 = -w * w;
    Mat3 r = {{1 + 2 * (t8 + t10), 2 * (t6 - t4), 2 * (t3 + t7),  //
               2 * (t4 + t6), 1 + 2 * (t5 + t10), 2 * (t9 - t2),  //
               2 * (t7 - t3), 2 * (t2 + t9), 1 + 2 * (t5 + t8)}};
    return r;
}
// (a, b) -> (a c - b s, a s + b c): the in-place rotation pattern of one_leg.cu:15-23,146-156
Mat3 rot_xy(double c, double s) { return {{c, -s, 0, s, c, 0, 0, 0, 1}}; }
Mat3 rot_xz(double c, double s) { return {{c, 0, -s, 0, 1, 0, s, 0, c}}; }

AngleTest make_angle_gt(double theta) {
    AngleTest t{};
    if (theta >= M_PI) {  // atan2f never exceeds pi (float pi < M_PI): always false
        t.c = 0, t.ns = 0, t.bias = -1, t.lower = 0;
    } else if (theta < -M_PI) {  // always true
        t.c = 0, t.ns = 0, t.bias = 1, t.lower = 1;
    } else {
        double c = std::cos(theta), s = std::sin(theta);
        // theta == 0: Y = +0, X < 0 (angle = pi) must count as "greater"; a denormal-scale
        // sine keeps the cross product strictly positive there without moving the threshold.
        if (theta == 0) s = 1e-30;
        t.c = (float)c;
        t.ns = (float)(-s);
        t.bias = 0;
        t.lower = theta < 0 ? 1u : 0u;
    }
    t.thr_dn = t.lower ? 0.f : kAngleBig;
    return t;
}
AngleTest make_angle_lt(double theta) { return make_angle_gt(-theta); }  // apply to (X, -Y)

PlanCircle make_circle(float cx, float cy, float r, bool attractive) {
    PlanCircle c;
    c.cx = cx, c.cy = cy, c.r = r;
    if (attractive) {
        double thr = (double)r + kCircleMargin;
        c.sgn = 1.f;
        c.thr_s = (float)(thr * thr);
    } else {
        double thr = (double)r - kCircleMargin;
        c.sgn = -1.f;
        c.thr_s = thr > 0 ? (float)(-(thr * thr)) : INFINITY;  // radius below the margin: always valid
    }
    return c;
}

// leg_geometry.cu.h:12-26
float min_reach(const lrm_leg_t& l) {
    const float x = l.femur_length + l.tibia_length * std::cos(l.min_angle_tibia);
    const float y = l.tibia_length * std::sin(l.min_angle_tibia);
    return std::sqrt(x * x + y * y);
}

void fill_planar(const lrm_leg_t& l, LegPlan* p) {
    p->coxa_length = l.coxa_length;

    // coxa limits
    const double cmax = l.max_angle_coxa, cmin = l.min_angle_coxa;
    p->over = make_angle_gt(cmax);
    p->under = make_angle_lt(cmin);
    p->mega_hi = make_angle_gt((double)(l.max_angle_coxa + kPi / 2));
    p->mega_lo = make_angle_lt((double)(l.min_angle_coxa - kPi / 2));
    p->mid = make_angle_gt((double)((l.max_angle_coxa + l.min_angle_coxa) / 2));
    p->cos_max = (float)std::cos(cmax), p->sin_max = (float)std::sin(cmax);
    p->cos_min = (float)std::cos(cmin), p->sin_min = (float)std::sin(cmin);

    // sectors, circles.cu.h:48-78
    const float middle = (std::max(l.tibia_absolute_neg, l.min_angle_femur) +
                          std::min(l.tibia_absolute_pos, l.max_angle_femur)) / 2;
    p->middle = make_angle_gt((double)middle);

    const float r_in = min_reach(l);
    const float r_out = l.tibia_length + l.femur_length;
    p->inner = make_circle(0.f, 0.f, r_in, false);
    p->outer = make_circle(0.f, 0.f, r_out, true);

    struct Raw {
        float cx, cy, r;
    };
    auto above = [&](float lim) {
        return Raw{l.tibia_length * std::cos(lim), l.tibia_length * std::sin(lim), l.femur_length};
    };
    auto winglet = [&](bool lower_side) {
        const float a = lower_side ? l.min_angle_femur : l.max_angle_femur;
        return Raw{std::cos(a) * l.femur_length, std::sin(a) * l.femur_length, l.tibia_length};
    };

    for (int u = 0; u < 2; u++) {
        const bool upper = u != 0;
        const float fem_lim = upper ? l.max_angle_femur : l.min_angle_femur;
        const float abs_lim = upper ? l.tibia_absolute_pos : l.tibia_absolute_neg;
        const float fem_lim_o = !upper ? l.max_angle_femur : l.min_angle_femur;
        const float abs_lim_o = !upper ? l.tibia_absolute_pos : l.tibia_absolute_neg;
        const bool fem_first = (!upper) ^ (fem_lim < abs_lim);
        const bool fem_first_o = (!upper) ^ (fem_lim_o < abs_lim_o);
        p->sat[u] = make_angle_gt((double)(fem_first ? fem_lim : abs_lim));

        // circles.cu.h:337-383
        Raw t[3] = {above(l.tibia_absolute_neg), above(l.tibia_absolute_pos), winglet(!upper)};
        bool att[3] = {false, false, false};
        const int excl = upper ? 0 : 1;
        if (fem_first_o) t[excl] = winglet(upper);  // the other side's winglet blocks instead
        const int other = upper ? 1 : 0;
        att[other] = !fem_first;
        att[2] = fem_first;
        p->att_slot[u] = att[other] ? other : 2;
        for (int j = 0; j < 3; j++) p->slot[u][j] = make_circle(t[j].cx, t[j].cy, t[j].r, att[j]);
    }

    // corner points, circles.cu.h:417-476
    float fem[10], tib[10];
    fem[0] = l.min_angle_femur, tib[0] = l.max_angle_tibia;
    fem[1] = l.min_angle_femur, tib[1] = l.min_angle_tibia;
    fem[2] = l.min_angle_femur, tib[2] = l.tibia_absolute_neg - fem[2];
    fem[3] = l.tibia_absolute_neg - l.min_angle_tibia, tib[3] = l.tibia_absolute_neg - fem[3];
    fem[4] = l.tibia_absolute_neg - l.max_angle_tibia, tib[4] = l.tibia_absolute_neg - fem[4];
    fem[5] = l.max_angle_femur, tib[5] = l.min_angle_tibia;
    fem[6] = l.max_angle_femur, tib[6] = l.max_angle_tibia;
    fem[7] = l.max_angle_femur, tib[7] = l.tibia_absolute_pos - fem[7];
    fem[8] = l.tibia_absolute_pos - l.min_angle_tibia, tib[8] = l.tibia_absolute_pos - fem[8];
    fem[9] = fem[8], tib[9] = tib[8];  // the reference emits this pair twice (circles.cu.h:447-450)
    p->n_corners = 0;
    for (int i = 0; i < kMaxCorners; i++) p->corner_x[i] = p->corner_y[i] = 0.f;
    for (int i = 0; i < 10; i++) {
        const float f = fem[i], tb = tib[i], a = f + tb;
        const bool ok = (double)f < (double)l.max_angle_femur + kCornerEps &&
                        (double)f > (double)l.min_angle_femur - kCornerEps &&
                        (double)tb < (double)l.max_angle_tibia + kCornerEps &&
                        (double)tb > (double)l.min_angle_tibia - kCornerEps &&
                        (double)a < (double)l.tibia_absolute_pos + kCornerEps &&
                        (double)a > (double)l.tibia_absolute_neg - kCornerEps;
        if (!ok) continue;
        const float xf = l.femur_length * std::cos(f), yf = l.femur_length * std::sin(f);
        const float xt = l.tibia_length * std::cos(a), yt = l.tibia_length * std::sin(a);
        p->corner_x[p->n_corners] = xf + xt;
        p->corner_y[p->n_corners] = yf + yt;
        p->n_corners++;
    }
}

// ---- valid arcs -------------------------------------------------------------------------------
// For circle j of a sector: which directions phi around its centre put the point
// c_j + r_j (cos phi, sin phi) inside every OTHER constraint of the sector?  Breakpoints are the
// intersections with the other circles at their margin-adjusted radii; each elementary interval is
// classified by testing its midpoint.  Returns the number of maximal valid arcs and the first one.
struct Arc {
    double lo, hi;  // hi > lo, hi - lo <= 2 pi
};
bool constraint_ok(const PlanCircle& k, double x, double y) {
    const double dx = x - k.cx, dy = y - k.cy, m = std::sqrt(dx * dx + dy * dy);
    return k.sgn > 0 ? (m < (double)k.r + kCircleMargin) : (m > (double)k.r - kCircleMargin);
}
int valid_arcs(const PlanCircle* set, int j, Arc* first /* room for 2 */) {
    const PlanCircle& c = set[j];
    std::vector<double> brk;
    for (int k = 0; k < 4; k++) {
        if (k == j) continue;
        const double R = (double)set[k].r + (set[k].sgn > 0 ? kCircleMargin : -kCircleMargin);
        const double dx = (double)set[k].cx - c.cx, dy = (double)set[k].cy - c.cy;
        const double d = std::sqrt(dx * dx + dy * dy), r = c.r;
        if (d <= 0 || R <= 0 || r <= 0) continue;
        const double cosb = (r * r + d * d - R * R) / (2 * r * d);
        if (cosb <= -1 || cosb >= 1) continue;
        const double a = std::atan2(dy, dx), b = std::acos(cosb);
        brk.push_back(a + b);
        brk.push_back(a - b);
    }
    auto ok_at = [&](double phi) {
        const double x = c.cx + (double)c.r * std::cos(phi), y = c.cy + (double)c.r * std::sin(phi);
        for (int k = 0; k < 4; k++)
            if (k != j && !constraint_ok(set[k], x, y)) return false;
        return true;
    };
    if (brk.empty()) {
        if (!ok_at(0.0)) return 0;
        first->lo = 0, first->hi = 2 * M_PI;
        return 1;
    }
    for (double& b : brk) {
        b = std::fmod(b, 2 * M_PI);
        if (b < 0) b += 2 * M_PI;
    }
    std::sort(brk.begin(), brk.end());
    const size_t n = brk.size();
    std::vector<char> good(n);
    for (size_t i = 0; i < n; i++) {
        const double lo = brk[i], hi = (i + 1 < n) ? brk[i + 1] : brk[0] + 2 * M_PI;
        good[i] = ok_at(0.5 * (lo + hi)) ? 1 : 0;
    }
    // merge circularly: count starts of valid runs
    int arcs = 0;
    bool all = true;
    for (size_t i = 0; i < n; i++) all = all && good[i];
    if (all) {
        first->lo = 0, first->hi = 2 * M_PI;
        return 1;
    }
    for (size_t i = 0; i < n; i++) {
        if (good[i] && !good[(i + n - 1) % n]) {
            if (arcs < 2) {
                size_t e = i;
                while (good[(e + 1) % n]) e = (e + 1) % n;
                first[arcs].lo = brk[i];
                first[arcs].hi = brk[(e + 1) % n];
                if (first[arcs].hi <= first[arcs].lo) first[arcs].hi += 2 * M_PI;
            }
            arcs++;
        }
    }
    return arcs;
}

void fill_sector_rows(LegPlan* p) {
    p->generic = 0;
    for (int s = 0; s < 4; s++) {
        const int upper = s >> 1, ext = s & 1;
        PlanCircle set[4];
        set[0] = p->inner;
        for (int j = 0; j < 3; j++) {
            set[j + 1] = p->slot[upper][j];
            if (ext && p->att_slot[upper] == j) set[j + 1] = p->outer;
        }
        // arcs: [circle][0..1] -> (ax, ay, ah); only the inner circle may need two
        float arc[4][2][3];
        for (int j = 0; j < 4; j++) {
            Arc a[2] = {{0, 0}, {0, 0}};
            const int n = valid_arcs(set, j, a);
            if (n > (j == 0 ? 2 : 1)) p->generic = 1;
            for (int k = 0; k < 2; k++) {
                if (k >= n) {
                    arc[j][k][0] = 1.f, arc[j][k][1] = 0.f, arc[j][k][2] = 2.f;  // empty
                    continue;
                }
                const double mid = 0.5 * (a[k].lo + a[k].hi), half = 0.5 * (a[k].hi - a[k].lo);
                arc[j][k][0] = (float)std::cos(mid), arc[j][k][1] = (float)std::sin(mid);
                arc[j][k][2] = half >= M_PI ? -2.f : (float)std::cos(half);
            }
        }
        for (int j = 0; j < 3; j++) {
            float* o = p->sector[s].slot[j];
            const PlanCircle& c = set[j + 1];
            o[0] = c.cx, o[1] = c.cy, o[2] = c.r, o[3] = c.sgn;
            o[4] = arc[j + 1][0][0], o[5] = arc[j + 1][0][1], o[6] = arc[j + 1][0][2];
            o[7] = arc[0][0][j];  // inner circle's first arc rides in the spare column
        }
        for (int k = 0; k < 3; k++) p->sector[s].inner_b[k] = arc[0][1][k];
        p->sector[s].inner_b[3] = 0.f;
    }
    auto plain = [](const AngleTest& t) { return t.bias == 0.f && t.lower == 0; };
    p->std_coxa = plain(p->over) && plain(p->under) && plain(p->mega_hi) && plain(p->mega_lo) &&
                  plain(p->mid);
}

void store(const Mat3& m, float* out) {
    for (int i = 0; i < 9; i++) out[i] = (float)m.m[i];
}
}  // namespace

// ---- quaternion helpers (reference layouts) -------------------------------------------------
void quat_from_vect_angle(const float axis[3], float angle, float out[4]) {
    float s, c;
    sincosf(angle / 2, &s, &c);
    const float mag = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    out[0] = s;
    out[1] = c * axis[0] / mag;
    out[2] = c * axis[1] / mag;
    out[3] = c * axis[2] / mag;
}
void quat_multiply(const float a[4], const float b[4], float out[4]) {
    const float ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const float bx = b[0], by = b[1], bz = b[2], bw = b[3];
    const float w = aw * bw - ax * bx - ay * by - az * bz;
    const float x = aw * bx + ax * bw + ay * bz - az * by;
    const float y = aw * by - ax * bz + ay * bw + az * bx;
    const float z = aw * bz + ax * by - ay * bx + az * bw;
    out[0] = x, out[1] = y, out[2] = z, out[3] = w;
}
void quat_invert(const float q[4], float out[4]) {
    const float n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    out[0] = q[0] / n, out[1] = -q[1] / n, out[2] = -q[2] / n, out[3] = -q[3] / n;
}
void quat_rotate(const float q[4], const float v[3], float out[3]) {
    const Mat3 m = quat_matrix(q);
    for (int i = 0; i < 3; i++)
        out[i] = (float)(m.m[3 * i] * v[0] + m.m[3 * i + 1] * v[1] + m.m[3 * i + 2] * v[2]);
}
// one_leg_global.cu:48-60 / several_leg.cu:743-754 with rpyFromQuat's pitch
// (unified_math_cuda.cu.h:71-76: float products widened to double, double asin).
float quat_pitch_for_leg(const float quat[4], float body_angle) {
    const float az[3] = {0, 0, 1};
    float qa[4], qai[4], tmp[4], res[4];
    quat_from_vect_angle(az, body_angle, qa);
    quat_invert(qa, qai);
    quat_multiply(qa, quat, tmp);
    quat_multiply(tmp, qai, res);
    const float x = res[0], y = res[1], z = res[2], w = res[3];
    const double sinp = 2 * (w * y - z * x);
    if (std::fabs(sinp) >= 1) return copysignf((float)(M_PI / 2), (float)sinp);
    return (float)std::asin(sinp);
}

void default_leg(int robot, float azimuth, lrm_leg_t* out) {
    // static_variables.cpp:44-93 through leg_factory (:6-42)
    const float pitch_deg = robot == 0 ? 0.f : -45.f;
    const float tip = robot == 0 ? 160.f : 135.f;
    lrm_leg_t l;
    std::memset(&l, 0, sizeof l);
    l.coxa_pitch = pitch_deg / 180.f * kPi;
    l.body = 181.f;
    l.coxa_length = 65.5f;
    l.femur_length = 129.f;
    l.tibia_length = tip;
    l.tibia_absolute_pos = -5.f / 180.0f * kPi - l.coxa_pitch;
    l.tibia_absolute_neg = (-180.0f - -5.f) / 180.0f * kPi - l.coxa_pitch;
    l.max_angle_coxa = kPi / 180.0f * 60.f;
    l.min_angle_coxa = -kPi / 180.0f * 60.f;
    l.max_angle_femur = kPi / 180.0f * 90.f;
    l.min_angle_femur = -kPi / 180.0f * 90.f;
    l.max_angle_tibia = kPi / 180.0f * 120.f;
    l.min_angle_tibia = -kPi / 180.0f * 120.f;
    l.body_angle = azimuth;
    *out = l;
}

// ---- plans -----------------------------------------------------------------------------------
static void build_common(const lrm_leg_t& leg, const float* quat, bool points_in_world,
                         LegPlan* out) {
    std::memset(out, 0, sizeof *out);
    const float ident[4] = {1.f, 0.f, 0.f, 0.f};
    const float* q = quat ? quat : ident;

    lrm_leg_t o = leg;  // rotate_leg_data: only the absolute tibia limits move
    const float pitch = quat_pitch_for_leg(q, leg.body_angle);
    o.tibia_absolute_pos -= pitch;
    o.tibia_absolute_neg -= pitch;
    fill_planar(o, out);
    fill_sector_rows(out);

    float qi[4];
    quat_invert(q, qi);
    float s_az, c_az, s_p, c_p;
    sincosf(-o.body_angle, &s_az, &c_az);
    sincosf(-o.coxa_pitch, &s_p, &c_p);
    const Mat3 to_body = quat_matrix(qi);
    const Mat3 az = rot_xy(c_az, s_az);
    const Mat3 pit = rot_xz(c_p, s_p);
    const Mat3 fwd = points_in_world ? mul(pit, mul(az, to_body)) : mul(pit, az);
    store(fwd, out->M);
    out->t[0] = (float)(pit.m[0] * -(double)o.body);
    out->t[1] = (float)(pit.m[3] * -(double)o.body);
    out->t[2] = (float)(pit.m[6] * -(double)o.body);

    float s_pr, c_pr;
    sincosf(o.coxa_pitch, &s_pr, &c_pr);
    const Mat3 pit_back = rot_xz(c_pr, s_pr);
    const Mat3 az_back = rot_xy(c_az, -s_az);  // z_unrotateInPlace, one_leg_global.cu:33-39
    const Mat3 back = points_in_world ? mul(quat_matrix(q), mul(az_back, pit_back))
                                      : mul(az_back, pit_back);
    store(back, out->Mo);

    // gravity-side test of reachable_rotate_leg (several_leg.cu:58-62): x component of
    // Rz(-body_angle) * qtRotate(qtInvert(q), v)
    const Mat3 g = mul(az, to_body);
    out->grav[0] = (float)g.m[0], out->grav[1] = (float)g.m[1], out->grav[2] = (float)g.m[2];
}

void build_leg_plan(const lrm_leg_t& leg, const float* quat, LegPlan* out) {
    build_common(leg, quat, true, out);
}
void build_leg_plan_rotated_limits(const lrm_leg_t& leg, const float* quat, LegPlan* out) {
    build_common(leg, quat, false, out);
}

}  // namespace lrm
