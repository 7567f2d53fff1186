/* TEST INFRASTRUCTURE ONLY — plain-C CPU restatement of the reference hot path.
 *
 * What this is: an independent re-expression, in C11, of the algorithm the reference implements
 * in CUDA `__host__ __device__` code, following the reference's HOST arithmetic operation by
 * operation (float ops in the same order, double where the reference's literals promote to
 * double, glibc libm, no FMA contraction: build with -ffp-contract=off).  Each function cites the
 * reference file:line it follows.  It is pinned (bit-exact) against the compiled reference
 * (oracle/_ref/libref_oracle.so, built from /root/reference by oracle/Makefile) by
 * tests/test_oracle_golden.py in the build container, and against the committed golden vectors in
 * tests/golden/ (generated from the compiled reference by tests/golden/make_golden.py) anywhere.
 *
 * The multi-leg part (op_standability) restates code that is __device__-only in the reference
 * (several_leg.cu) and therefore has no compiled host twin; its leaf predicates are pinned
 * through the one-leg functions and the quaternion helpers it is built from, and its pipeline
 * logic is "parity unpinned" against a running reference (no GPU here; DESIGN.md says so).
 *
 * Who may use it: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
 * The product never links this file.
 */
#define _GNU_SOURCE
#include "oracle_port.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* settings.h:9, circles.cu.h:7 — both are *double* literals in the reference; comparisons
 * against them promote the float operand to double. */
#define OP_CIRCLE_MARGIN 0.001
#define OP_EPS 0.001
static const float OP_PI = 3.14159265358979323846264338327950288419716939937510582097f;

typedef struct { float x, y, radius; int attractive; } op_circle;

/* ------------------------------------------------------------------------------------------ */
/* static_variables.cpp:6-42 — note femur_length <- tibia2femur, tibia_length <- femur2tip     */
static op_leg_t op_leg_factory(float azimuth, float body2coxa, float coxa_pitch_deg,
                               float coxa2tibia, float tibia2femur, float femur2tip,
                               float coxa_deg, float femur_deg, float tibia_deg, float tib_abs_pos,
                               float tib_abs_neg) {
    op_leg_t l;
    memset(&l, 0, sizeof l);
    l.coxa_pitch = coxa_pitch_deg / 180.f * OP_PI;
    l.body = body2coxa;
    l.coxa_length = coxa2tibia;
    l.femur_length = tibia2femur;
    l.tibia_length = femur2tip;
    l.tibia_absolute_pos = tib_abs_pos / 180.0f * OP_PI - l.coxa_pitch;
    l.tibia_absolute_neg = (-180.0f - tib_abs_neg) / 180.0f * OP_PI - l.coxa_pitch;
    l.max_angle_coxa = OP_PI / 180.0f * coxa_deg;
    l.min_angle_coxa = -OP_PI / 180.0f * coxa_deg;
    l.max_angle_femur = OP_PI / 180.0f * femur_deg;
    l.min_angle_femur = -OP_PI / 180.0f * femur_deg;
    l.max_angle_tibia = OP_PI / 180.0f * tibia_deg;
    l.min_angle_tibia = -OP_PI / 180.0f * tibia_deg;
    l.body_angle = azimuth;
    return l;
}

/* static_variables.cpp:44-93 */
void op_get_leg(int robot, float azimuth, op_leg_t* out) {
    if (robot == 0)
        *out = op_leg_factory(azimuth, 181.f, 0.f, 65.5f, 129.f, 160.f, 60.f, 90.f, 120.f, -5.f, -5.f);
    else
        *out = op_leg_factory(azimuth, 181.f, -45.f, 65.5f, 129.f, 135.f, 60.f, 90.f, 120.f, -5.f, -5.f);
}

/* ------------------------------------------------------------------------------------------ */
/* Quaternion helpers, unified_math_cuda.cu.h:13-83.  Layouts are deliberately the reference's. */
op_f3 op_qt_rotate(op_f4 q, op_f3 v) { /* :13-27 — .x plays the scalar part */
    float t2 = q.x * q.y, t3 = q.x * q.z, t4 = q.x * q.w;
    float t5 = -q.y * q.y, t6 = q.y * q.z, t7 = q.y * q.w;
    float t8 = -q.z * q.z, t9 = q.z * q.w, t10 = -q.w * q.w;
    op_f3 r;
    r.x = 2.0f * ((t8 + t10) * v.x + (t6 - t4) * v.y + (t3 + t7) * v.z) + v.x;
    r.y = 2.0f * ((t4 + t6) * v.x + (t5 + t10) * v.y + (t9 - t2) * v.z) + v.y;
    r.z = 2.0f * ((t7 - t3) * v.x + (t2 + t9) * v.y + (t5 + t8) * v.z) + v.z;
    return r;
}
op_f4 op_qt_invert(op_f4 q) { /* :29-34 */
    float n = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    op_f4 r = {q.x / n, -q.y / n, -q.z / n, -q.w / n};
    return r;
}
op_f4 op_qt_multiply(op_f4 a, op_f4 b) { /* :40-46 — .w plays the scalar part */
    float w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    float x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    float y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    float z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    op_f4 r = {x, y, z, w};
    return r;
}
op_f4 op_quat_from_vect_angle(op_f3 axis, float angle) { /* :48-57 */
    float s, c;
    sincosf(angle / 2, &s, &c);
    float mag = sqrtf(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
    op_f4 r = {s, c * axis.x / mag, c * axis.y / mag, c * axis.z / mag};
    return r;
}
op_f3 op_rpy_from_quat(op_f4 q) { /* :59-83 — float products widened to double, double libm */
    const float x = q.x, y = q.y, z = q.z, w = q.w;
    op_f3 rpy;
    double sinr_cosp = 2 * (w * x + y * z);
    double cosr_cosp = 1 - 2 * (x * x + y * y);
    rpy.x = (float)atan2(sinr_cosp, cosr_cosp);
    double sinp = 2 * (w * y - z * x);
    if (fabs(sinp) >= 1)
        rpy.y = copysignf((float)(M_PI / 2), (float)sinp);
    else
        rpy.y = (float)asin(sinp);
    double siny_cosp = 2 * (w * z + x * y);
    double cosy_cosp = 1 - 2 * (y * y + z * z);
    rpy.z = (float)atan2(siny_cosp, cosy_cosp);
    return rpy;
}
op_f4 op_rpy_to_quat(float r, float p, float y) { /* octree_util.cu.h:164-172 */
    op_f3 ax = {1, 0, 0}, ay = {0, 1, 0}, az = {0, 0, 1};
    op_f4 qr = op_quat_from_vect_angle(ax, r);
    op_f4 qp = op_qt_multiply(op_quat_from_vect_angle(ay, p), qr);
    return op_qt_multiply(op_quat_from_vect_angle(az, y), qp);
}
/* one_leg_global.cu:48-60 (== several_leg.cu:743-754): body orientation shifts the absolute
 * tibia limits by the pitch seen from the leg's azimuth. */
op_leg_t op_rotate_leg_data(op_f4 quat, op_leg_t leg) {
    op_f3 az = {0, 0, 1};
    op_f4 qa = op_quat_from_vect_angle(az, leg.body_angle);
    op_f4 res = op_qt_multiply(op_qt_multiply(qa, quat), op_qt_invert(qa));
    float pitch = op_rpy_from_quat(res).y;
    leg.tibia_absolute_pos -= pitch;
    leg.tibia_absolute_neg -= pitch;
    return leg;
}
void op_qt_rotate_p(const float* q, const float* v, float* o) {
    op_f4 qq = {q[0], q[1], q[2], q[3]};
    op_f3 vv = {v[0], v[1], v[2]};
    op_f3 r = op_qt_rotate(qq, vv);
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void op_qt_multiply_p(const float* a, const float* b, float* o) {
    op_f4 aa = {a[0], a[1], a[2], a[3]}, bb = {b[0], b[1], b[2], b[3]};
    op_f4 r = op_qt_multiply(aa, bb);
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
void op_quat_from_vect_angle_p(const float* ax, float angle, float* o) {
    op_f3 a = {ax[0], ax[1], ax[2]};
    op_f4 r = op_quat_from_vect_angle(a, angle);
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
void op_rpy_to_quat_p(float r, float p, float y, float* o) {
    op_f4 q = op_rpy_to_quat(r, p, y);
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}
void op_rotate_leg_data_p(const float* q, const op_leg_t* leg, op_leg_t* out) {
    op_f4 qq = {q[0], q[1], q[2], q[3]};
    *out = op_rotate_leg_data(qq, *leg);
}

/* ------------------------------------------------------------------------------------------ */
/* Planar geometry in the femur frame.                                                         */

/* leg_geometry.cu.h:12-26 (lower side): closest the gripper gets to the femur joint */
static float op_min_reach(const op_leg_t* l) {
    float x = l->femur_length + l->tibia_length * cosf(l->min_angle_tibia);
    float y = l->tibia_length * sinf(l->min_angle_tibia);
    return sqrtf(x * x + y * y);
}
/* circles.cu.h:80-122 */
static op_circle op_c_inner(const op_leg_t* l) { op_circle c = {0, 0, op_min_reach(l), 0}; return c; }
static op_circle op_c_outer(const op_leg_t* l) {
    op_circle c = {0, 0, l->tibia_length + l->femur_length, 1};
    return c;
}
static op_circle op_c_above(const op_leg_t* l, float abs_limit) {
    op_circle c = {l->tibia_length * cosf(abs_limit), l->tibia_length * sinf(abs_limit),
                   l->femur_length, 0};
    return c;
}
static op_circle op_c_winglet(const op_leg_t* l, int lower_side) { /* :116-122, leg_geometry :32-37 */
    float a = lower_side ? l->min_angle_femur : l->max_angle_femur;
    op_circle c = {cosf(a) * l->femur_length, sinf(a) * l->femur_length, l->tibia_length, 0};
    return c;
}

/* circles.cu.h:48-78.  bit0 Upper, bit1 FullyExtended, bit2 FemurLim, bit3 FemurLim_other */
int op_find_region(float x, float y, const op_leg_t* d) {
    float angle = atan2f(y, x);
    float middle = (fmaxf(d->tibia_absolute_neg, d->min_angle_femur) +
                    fminf(d->tibia_absolute_pos, d->max_angle_femur)) / 2;
    int upper = angle > middle;
    float fem_lim = upper ? d->max_angle_femur : d->min_angle_femur;
    float abs_lim = upper ? d->tibia_absolute_pos : d->tibia_absolute_neg;
    float fem_lim_o = !upper ? d->max_angle_femur : d->min_angle_femur;
    float abs_lim_o = !upper ? d->tibia_absolute_pos : d->tibia_absolute_neg;
    int fem_first = (!upper) ^ (fem_lim < abs_lim);
    int fem_first_o = (!upper) ^ (fem_lim_o < abs_lim_o);
    float sat = fem_first ? fem_lim : abs_lim;
    int extended = upper ^ (angle > sat);
    return upper | (extended << 1) | (fem_first << 2) | (fem_first_o << 3);
}

/* circles.cu.h:337-383 (MegaClamp == 0): slot 0 inner (repulsive), slots 1/2 the negative/positive
 * side circles, slot 3 this side's winglet; the attractive one becomes the outer circle when the
 * leg is fully extended on this side. */
static int op_build_circles(const op_leg_t* l, int region, op_circle* c) {
    const int upper = region & 1, extended = (region >> 1) & 1;
    const int fem_first = (region >> 2) & 1, fem_first_o = (region >> 3) & 1;
    const int lower_side = !upper;
    c[0] = op_c_inner(l);
    op_circle* t = c + 1;
    t[0] = op_c_above(l, l->tibia_absolute_neg);
    t[1] = op_c_above(l, l->tibia_absolute_pos);
    int excl = upper ? 0 : 1;
    if (fem_first_o) t[excl] = op_c_winglet(l, !lower_side);
    t[excl].attractive = 0;
    int other = !upper ? 0 : 1;
    t[2] = op_c_winglet(l, lower_side);
    t[other].attractive = !fem_first;
    t[2].attractive = fem_first;
    if (extended) {
        int idx = t[other].attractive ? other : 2;
        t[idx] = op_c_outer(l);
    }
    return 4;
}
int op_insert_circles(float x, float y, const op_leg_t* leg, float* out16) {
    op_circle c[4];
    int n = op_build_circles(leg, op_find_region(x, y, leg), c);
    for (int i = 0; i < n; i++) {
        out16[4 * i] = c[i].x; out16[4 * i + 1] = c[i].y;
        out16[4 * i + 2] = c[i].radius; out16[4 * i + 3] = (float)c[i].attractive;
    }
    return n;
}

/* circles.cu.h:417-476: corner points of the planar workspace = FK at joint-limit combinations,
 * kept when all three limit pairs hold within EPS (double arithmetic for the +-EPS terms). */
static int op_build_corners(const op_leg_t* l, op_circle* out) {
    float fem[10], tib[10];
    fem[0] = l->min_angle_femur; tib[0] = l->max_angle_tibia;
    fem[1] = l->min_angle_femur; tib[1] = l->min_angle_tibia;
    fem[2] = l->min_angle_femur; tib[2] = l->tibia_absolute_neg - fem[2];
    fem[3] = l->tibia_absolute_neg - l->min_angle_tibia; tib[3] = l->tibia_absolute_neg - fem[3];
    fem[4] = l->tibia_absolute_neg - l->max_angle_tibia; tib[4] = l->tibia_absolute_neg - fem[4];
    fem[5] = l->max_angle_femur; tib[5] = l->min_angle_tibia;
    fem[6] = l->max_angle_femur; tib[6] = l->max_angle_tibia;
    fem[7] = l->max_angle_femur; tib[7] = l->tibia_absolute_pos - fem[7];
    fem[8] = l->tibia_absolute_pos - l->min_angle_tibia; tib[8] = l->tibia_absolute_pos - fem[8];
    fem[9] = l->tibia_absolute_pos - l->min_angle_tibia; tib[9] = l->tibia_absolute_pos - fem[9];
    int n = 0;
    for (int i = 0; i < 10; i++) {
        float f = fem[i], t = tib[i];
        int f_ok = ((double)f < (double)l->max_angle_femur + OP_EPS) &&
                   ((double)f > (double)l->min_angle_femur - OP_EPS);
        int t_ok = ((double)t < (double)l->max_angle_tibia + OP_EPS) &&
                   ((double)t > (double)l->min_angle_tibia - OP_EPS);
        float a = f + t;
        int a_ok = ((double)a < (double)l->tibia_absolute_pos + OP_EPS) &&
                   ((double)a > (double)l->tibia_absolute_neg - OP_EPS);
        if (f_ok && t_ok && a_ok) {
            float xf = l->femur_length * cosf(f), yf = l->femur_length * sinf(f);
            float xt = l->tibia_length * cosf(a), yt = l->tibia_length * sinf(a);
            op_circle p = {xf + xt, yf + yt, 0, 1};
            out[n++] = p;
        }
    }
    return n;
}
int op_insert_intersec(const op_leg_t* leg, float* out20) {
    op_circle c[10];
    int n = op_build_corners(leg, c);
    for (int i = 0; i < n; i++) { out20[2 * i] = c[i].x; out20[2 * i + 1] = c[i].y; }
    return n;
}

/* one_leg.cu:31-41 */
static int op_circle_ok(const op_circle* c, float x, float y, float* dist_out) {
    x -= c->x; y -= c->y;
    float m = sqrtf(x * x + y * y);
    float d = c->radius - m;
    int inside = !signbit(d);
    *dist_out = d;
    return (inside == c->attractive) || ((double)fabsf(d) < OP_CIRCLE_MARGIN);
}
/* one_leg.cu:65-89 with OnlyCircles = true */
static int op_all_circles_ok(float x, float y, const op_circle* c, int n) {
    for (int i = 0; i < n; i++) {
        float d;
        if (!op_circle_ok(&c[i], x, y, &d)) return 0;
    }
    return 1;
}
/* one_leg.cu:42-63: project (x,y) on the circle; a point at the centre projects along +x */
static int op_project(const op_circle* c, float* x, float* y, float* dist_out) {
    *x -= c->x; *y -= c->y;
    float m = sqrtf(*x * *x + *y * *y);
    float d = c->radius - m;
    int inside = !signbit(d);
    int ok = (inside == c->attractive) || ((double)fabsf(d) < OP_CIRCLE_MARGIN);
    if ((double)m < OP_CIRCLE_MARGIN) { *x = 1; *y = 0; m = 1; }
    float k = c->radius / m;
    *x = c->x + *x * k;
    *y = c->y + *y * k;
    *dist_out = d;
    return ok;
}
/* one_leg.cu:91-145 (CIRCLE_ARR_ORDERED, MegaClamp == 0): nearest valid boundary candidate.
 * Circles first (their projection must satisfy the 4 circles), then corner points, which only
 * compete when the query point itself is outside.  Ties keep the earlier candidate. */
static int op_clamp(float* x, float* y, const op_circle* c, int n) {
    int overall = 1;
    float bx = 0, by = 0, best = 999999999999999.9;
    for (int i = 0; i < n; i++) {
        float px = *x, py = *y, d;
        int ok = op_project(&c[i], &px, &py, &d);
        int cand_ok;
        if (fabsf(c[i].radius) < OP_CIRCLE_MARGIN) {
            if (overall) break;
            cand_ok = 1;
        } else {
            cand_ok = op_all_circles_ok(px, py, c, 4);
            overall = overall && ok;
        }
        if (cand_ok && fabsf(best) > fabsf(d)) { best = d; bx = px; by = py; }
    }
    *x -= bx;
    *y -= by;
    return overall;
}
/* one_leg.cu:167-208 */
static int op_plane_reach(float x, float y, const op_leg_t* d) {
    x -= d->coxa_length;
    op_circle c[4];
    int n = op_build_circles(d, op_find_region(x, y, d), c);
    return op_all_circles_ok(x, y, c, n);
}
static int op_plane_dist(float* x, float* y, const op_leg_t* d) {
    *x -= d->coxa_length;
    op_circle c[14];
    int n = op_build_circles(d, op_find_region(*x, *y, d), c);
    n += op_build_corners(d, c + n);
    return op_clamp(x, y, c, n);
}

/* ------------------------------------------------------------------------------------------ */
/* 3-D wrappers                                                                                */
static void op_rot2(float* a, float* b, float angle, float* c_out, float* s_out) {
    /* the (x*cos - y*sin, x*sin + y*cos) pattern of one_leg.cu:15-23,146-156 */
    float s, c;
    sincosf(angle, &s, &c);
    float buf = *a * s;
    *a = *a * c - *b * s;
    *b = buf + *b * c;
    if (c_out) { *c_out = c; *s_out = s; }
}
/* one_leg.cu:158-165 */
static void op_unrot_xy(op_f3* p, float c, float s) {
    float buf = p->y * s;
    p->y = -p->x * s + p->y * c;
    p->x = p->x * c + buf;
}
static float op_norm3(op_f3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); } /* host linorm */

/* one_leg.cu:280-319 */
int op_reachability_circles(op_f3 p, const op_leg_t* d) {
    p.x -= d->body;                                  /* place_over_coxa :9-24 */
    op_rot2(&p.x, &p.z, -d->coxa_pitch, NULL, NULL);
    float a;
    if (signbit(p.x)) a = atan2f(p.y * -1, p.x * -1); /* :290-303 — mirrored through the axis */
    else a = atan2f(p.y, p.x);
    if (a > d->max_angle_coxa || a < d->min_angle_coxa) return 0;
    op_rot2(&p.x, &p.y, -a, NULL, NULL);             /* cancel_coxa_rotation :146-156 */
    return op_plane_reach(p.x, p.z, d);
}

/* one_leg.cu:215-278 with Tout = bool */
static int op_finish_closest(op_f3* p, const op_leg_t* d, float a) {
    int mega = a > (d->max_angle_coxa + OP_PI / 2) || a < (d->min_angle_coxa - OP_PI / 2);
    float sat;
    if (mega) sat = (a > 0) ? a - OP_PI : a + OP_PI;
    else sat = fmaxf(fminf(a, d->max_angle_coxa), d->min_angle_coxa);
    int saturated = sat != a;
    float lim = (a > (d->max_angle_coxa + d->min_angle_coxa) / 2) ? d->max_angle_coxa
                                                                   : d->min_angle_coxa;
    float c1, s1;
    op_rot2(&p->x, &p->y, -sat, &c1, &s1);
    op_f3 save = *p;
    int valid = op_plane_dist(&p->x, &p->z, d);
    if (valid && !mega) {
        float c2, s2;
        op_rot2(&save.x, &save.y, -(lim - sat), &c2, &s2);
        save.x = 0;
        save.z = 0;
        if (op_norm3(*p) > op_norm3(save)) { /* the coxa-limit half-plane is closer */
            op_unrot_xy(&save, c2, s2);
            *p = save;
        }
    }
    op_unrot_xy(p, c1, s1);
    return valid && !saturated;
}
/* one_leg.cu:321-341 */
int op_distance_circles(op_f3* r, const op_leg_t* d) {
    op_f3 a = *r;
    a.x -= d->body;
    op_rot2(&a.x, &a.z, -d->coxa_pitch, NULL, NULL);
    op_f3 b = a;
    float ang = atan2f(a.y, a.x);
    float ang_flip = (ang > 0) ? ang - OP_PI : ang + OP_PI;
    int ra = op_finish_closest(&a, d, ang);
    int rb = op_finish_closest(&b, d, ang_flip);
    int direct = (!(ra ^ rb)) ? (op_norm3(a) < op_norm3(b)) : ra;
    *r = direct ? a : b;
    op_rot2(&r->x, &r->z, d->coxa_pitch, NULL, NULL); /* place_over_coxa<Reverse> */
    return ra || rb;
}

/* one_leg_global.cu:106-130, host branch */
int op_reachability_global(op_f3 p, const op_leg_t* dim, op_f4 quat) {
    op_leg_t o = op_rotate_leg_data(quat, *dim);
    op_f3 u = op_qt_rotate(op_qt_invert(quat), p);
    op_rot2(&u.x, &u.y, -o.body_angle, NULL, NULL);
    return op_reachability_circles(u, &o);
}
/* one_leg_global.cu:74-101, host branch */
int op_distance_global(op_f3* p, const op_leg_t* dim, op_f4 quat) {
    op_leg_t o = op_rotate_leg_data(quat, *dim);
    op_f3 u = op_qt_rotate(op_qt_invert(quat), *p);
    float c, s;
    op_rot2(&u.x, &u.y, -o.body_angle, &c, &s);
    int r = op_distance_circles(&u, &o);
    /* z_unrotateInPlace, one_leg_global.cu:33-39 */
    float buf = u.x * -s;
    u.x = u.x * c - u.y * -s;
    u.y = buf + u.y * c;
    *p = op_qt_rotate(quat, u);
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* slab loops (one_leg_global.cu:132-147) with an optional thread fan-out                      */
typedef struct {
    const float* xyz; size_t lo, hi; const op_leg_t* leg; op_f4 q;
    uint8_t* flag; float* out; int dist;
} op_job;
static void* op_worker(void* arg) {
    op_job* j = (op_job*)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        op_f3 p = {j->xyz[3 * i], j->xyz[3 * i + 1], j->xyz[3 * i + 2]};
        if (j->dist) {
            int f = op_distance_global(&p, j->leg, j->q);
            j->out[3 * i] = p.x; j->out[3 * i + 1] = p.y; j->out[3 * i + 2] = p.z;
            if (j->flag) j->flag[i] = (uint8_t)f;
        } else {
            j->flag[i] = (uint8_t)op_reachability_global(p, j->leg, j->q);
        }
    }
    return NULL;
}
static void op_run(op_job proto, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (n < 4096) threads = 1;
    pthread_t th[256];
    op_job jobs[256];
    size_t chunk = (n + (size_t)threads - 1) / (size_t)threads;
    int started = 0;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        jobs[t] = proto; jobs[t].lo = lo; jobs[t].hi = hi;
        if (threads == 1) { op_worker(&jobs[t]); return; }
        pthread_create(&th[t], NULL, op_worker, &jobs[t]);
        started++;
    }
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}
void op_reach(const float* xyz, size_t n, const op_leg_t* leg, const float* q, uint8_t* out,
              int threads) {
    op_job j = {xyz, 0, 0, leg, {q[0], q[1], q[2], q[3]}, out, NULL, 0};
    op_run(j, n, threads);
}
void op_dist(const float* xyz, size_t n, const op_leg_t* leg, const float* q, float* out_xyz,
             uint8_t* out_flag, int threads) {
    op_job j = {xyz, 0, 0, leg, {q[0], q[1], q[2], q[3]}, out_flag, out_xyz, 1};
    op_run(j, n, threads);
}

/* ------------------------------------------------------------------------------------------ */
/* Multi-leg positionability                                                                   */

/* several_leg.cu:811-857: roll in {-pi/8,0,pi/8} x pitch in {-pi/8,0,pi/8} x yaw in {0..pi/2 by
 * pi/8}; quat = yaw * (pitch * (roll * quatInit)) in the reference's multiply. */
int op_full_struct_orientations(float* out) {
    op_f3 ax = {1, 0, 0}, ay = {0, 1, 0}, az = {0, 0, 1};
    op_f4 q_init = op_quat_from_vect_angle(az, 0);
    const float r_min = -OP_PI / 8, r_max = OP_PI / 8, p_min = -OP_PI / 8, p_max = +OP_PI / 8;
    const float y_min = 0, y_max = OP_PI / 2;
    const int r_n = 2, p_n = 2, y_n = 4;
    int k = 0;
    for (int i = 0; i <= r_n; i++) {
        float roll = r_min + (r_max - r_min) * ((float)i / (float)r_n);
        op_f4 qr = op_qt_multiply(op_quat_from_vect_angle(ax, roll), q_init);
        for (int j = 0; j <= p_n; j++) {
            float pitch = p_min + (p_max - p_min) * ((float)j / (float)p_n);
            op_f4 qp = op_qt_multiply(op_quat_from_vect_angle(ay, pitch), qr);
            for (int m = 0; m <= y_n; m++) {
                float yaw = y_min + (y_max - y_min) * ((float)m / (float)y_n);
                op_f4 qy = op_qt_multiply(op_quat_from_vect_angle(az, yaw), qp);
                out[4 * k] = qy.x; out[4 * k + 1] = qy.y; out[4 * k + 2] = qy.z; out[4 * k + 3] = qy.w;
                k++;
            }
        }
    }
    return k;
}

/* octree_util.cu.h:184-198 (settings.h:35-38: 3x3x3 samples, roll +-pi/4, pitch/yaw +-pi/8).
 * The index remap (ind + ind/2) % 3 sends 2 -> 0, so only 8 distinct orientations exist. */
void op_quaternion_from_angle_index(unsigned idx, float* out4) {
    static const unsigned char samples[3] = {3, 3, 3};
    const float lim[6] = {-OP_PI / 4, OP_PI / 4, -OP_PI / 8, OP_PI / 8, -OP_PI / 8, OP_PI / 8};
    float rpy[3];
    unsigned rest = idx;
    for (int i = 0; i < 3; i++) {
        unsigned char n = samples[i];
        unsigned char ind = (unsigned char)(rest % n);
        ind = (unsigned char)((ind + (ind / 2)) % n);
        rest = rest / n;
        int den = n - 1 > 1 ? n - 1 : 1;
        float x = (float)ind / (unsigned char)den;
        rpy[i] = (1 - x) * lim[i * 2] + x * lim[i * 2 + 1];
    }
    op_rpy_to_quat_p(rpy[0], rpy[1], rpy[2], out4);
}

/* several_leg.cu:48-67 — target and body already rotated into the orientation frame */
static int op_leg_reaches(op_f3 target, op_f3 body, op_f4 q, const op_leg_t* dim) {
    target.x -= body.x; target.y -= body.y; target.z -= body.z;
    op_f3 g = op_qt_rotate(op_qt_invert(q), target);
    op_rot2(&g.x, &g.y, -dim->body_angle, NULL, NULL);
    if (g.x < 0) return 0;
    op_rot2(&target.x, &target.y, -dim->body_angle, NULL, NULL);
    return op_reachability_circles(target, dim);
}
/* collision.cu.h:5-23; the device uses norm3df, this host restatement sqrtf */
static int op_in_sphere(float radius, op_f3 c, op_f3 t) {
    float dx = c.x - t.x, dy = c.y - t.y, dz = c.z - t.z;
    return sqrtf(dx * dx + dy * dy + dz * dz) < radius;
}
static int op_in_cylinder(float radius, float plus_z, float minus_z, op_f3 c, op_f3 t) {
    float dz = t.z - c.z;
    float dx = t.x - c.x, dy = t.y - c.y;
    return (sqrtf(dx * dx + dy * dy + 0.f * 0.f) < radius) && (dz < plus_z) && (dz > minus_z);
}

/* xy bucket grid: an exactness-preserving accelerator for the "exists a target" reductions
 * (a target farther than `reach` in xy can satisfy none of the predicates). */
typedef struct {
    float x0, y0, inv; int nx, ny; int* start; int* idx;
} op_grid;
static void op_grid_build(op_grid* g, const op_f3* pts, size_t n, float cell) {
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (size_t i = 0; i < n; i++) {
        if (pts[i].x < xmin) xmin = pts[i].x;
        if (pts[i].x > xmax) xmax = pts[i].x;
        if (pts[i].y < ymin) ymin = pts[i].y;
        if (pts[i].y > ymax) ymax = pts[i].y;
    }
    if (n == 0) { xmin = ymin = 0; xmax = ymax = 1; }
    g->x0 = xmin; g->y0 = ymin; g->inv = 1.0f / cell;
    g->nx = (int)((xmax - xmin) * g->inv) + 1;
    g->ny = (int)((ymax - ymin) * g->inv) + 1;
    size_t cells = (size_t)g->nx * (size_t)g->ny;
    g->start = (int*)calloc(cells + 1, sizeof(int));
    g->idx = (int*)malloc((n ? n : 1) * sizeof(int));
    for (size_t i = 0; i < n; i++) {
        int cx = (int)((pts[i].x - xmin) * g->inv), cy = (int)((pts[i].y - ymin) * g->inv);
        g->start[(size_t)cy * g->nx + cx + 1]++;
    }
    for (size_t c = 0; c < cells; c++) g->start[c + 1] += g->start[c];
    int* fill = (int*)malloc(cells * sizeof(int));
    memcpy(fill, g->start, cells * sizeof(int));
    for (size_t i = 0; i < n; i++) {
        int cx = (int)((pts[i].x - xmin) * g->inv), cy = (int)((pts[i].y - ymin) * g->inv);
        g->idx[fill[(size_t)cy * g->nx + cx]++] = (int)i;
    }
    free(fill);
}
static void op_grid_free(op_grid* g) { free(g->start); free(g->idx); }
static void op_grid_range(const op_grid* g, float x, float y, float r, int* cx0, int* cx1, int* cy0,
                          int* cy1) {
    float fx0 = floorf((x - r - g->x0) * g->inv), fx1 = floorf((x + r - g->x0) * g->inv);
    float fy0 = floorf((y - r - g->y0) * g->inv), fy1 = floorf((y + r - g->y0) * g->inv);
    *cx0 = fx0 < 0 ? 0 : (int)fx0; *cy0 = fy0 < 0 ? 0 : (int)fy0;
    *cx1 = fx1 >= g->nx ? g->nx - 1 : (int)fx1; *cy1 = fy1 >= g->ny ? g->ny - 1 : (int)fy1;
}

typedef struct {
    /* shared, read-only */
    const op_f3* body_rot; const op_f3* targ_rot; const op_grid* grid;
    const op_leg_t* legs; int nlegs; op_f4 q; float reach;
    float radius_in, plus_in, minus_in, radius_out;
    const int* todo; size_t lo, hi; uint8_t* standable; uint8_t mark;
} op_pose_job;

/* one orientation of several_leg.cu:762-787 for a slab of still-unresolved bodies */
static void* op_pose_worker(void* arg) {
    op_pose_job* j = (op_pose_job*)arg;
    for (size_t k = j->lo; k < j->hi; k++) {
        int b = j->todo[k];
        op_f3 body = j->body_rot[b];
        int cx0, cx1, cy0, cy1;
        op_grid_range(j->grid, body.x, body.y, j->reach, &cx0, &cx1, &cy0, &cy1);
        /* eliminateFarAndColliding :504-559: some target inside the reach cylinder and none inside
         * the body cylinder (r = dim.body, z in (-110, 250)) */
        int near = 0, hit = 0;
        for (int cy = cy0; cy <= cy1 && !hit; cy++)
            for (int cx = cx0; cx <= cx1 && !hit; cx++) {
                size_t c = (size_t)cy * j->grid->nx + cx;
                for (int s = j->grid->start[c]; s < j->grid->start[c + 1]; s++) {
                    op_f3 t = j->targ_rot[j->grid->idx[s]];
                    if (!near && op_in_cylinder(j->radius_in, j->plus_in, j->minus_in, body, t)) near = 1;
                    if (op_in_cylinder(j->radius_out, 250.f, -110.f, body, t)) { hit = 1; break; }
                }
            }
        if (!near || hit) continue;
        /* eliminateUnreachable :633-706: every leg must reach at least one target */
        int all = 1;
        for (int l = 0; l < j->nlegs && all; l++) {
            int found = 0;
            for (int cy = cy0; cy <= cy1 && !found; cy++)
                for (int cx = cx0; cx <= cx1 && !found; cx++) {
                    size_t c = (size_t)cy * j->grid->nx + cx;
                    for (int s = j->grid->start[c]; s < j->grid->start[c + 1]; s++) {
                        if (op_leg_reaches(j->targ_rot[j->grid->idx[s]], body, j->q, &j->legs[l])) {
                            found = 1;
                            break;
                        }
                    }
                }
            all = found;
        }
        if (all) j->standable[b] = j->mark;
    }
    return NULL;
}

static float op_max_reach(const op_leg_t* legs, int nlegs) {
    float r = 0;
    for (int l = 0; l < nlegs; l++) {
        float v = fabsf(legs[l].body) + fabsf(legs[l].coxa_length) + fabsf(legs[l].femur_length) +
                  fabsf(legs[l].tibia_length);
        if (v > r) r = v;
    }
    return r + 1.0f;
}

void op_standability(const float* bodies, size_t nb, const float* targets, size_t nt,
                     const op_leg_t* legs, int nlegs, const float* quats, int nq, int pre_cull,
                     uint8_t* standable, int threads) {
    memset(standable, 0, nb);
    if (nb == 0 || nlegs <= 0) return;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    const op_f3* B = (const op_f3*)bodies;
    const op_f3* T = (const op_f3*)targets;
    uint8_t* body_alive = (uint8_t*)malloc(nb);
    memset(body_alive, 1, nb);
    op_f3* targ = (op_f3*)malloc((nt ? nt : 1) * sizeof(op_f3));
    size_t ntk = 0;

    if (pre_cull) {
        /* constructor :371-374 */
        op_grid g;
        op_grid_build(&g, T, nt, 100.f);
        for (size_t b = 0; b < nb; b++) {
            int cx0, cx1, cy0, cy1, collide = 0, close = 0;
            op_grid_range(&g, B[b].x, B[b].y, 401.f, &cx0, &cx1, &cy0, &cy1);
            for (int cy = cy0; cy <= cy1 && !collide; cy++)
                for (int cx = cx0; cx <= cx1 && !collide; cx++) {
                    size_t c = (size_t)cy * g.nx + cx;
                    for (int s = g.start[c]; s < g.start[c + 1]; s++) {
                        op_f3 t = T[g.idx[s]];
                        if (op_in_sphere(60.f, B[b], t)) { collide = 1; break; }  /* :413-440 */
                        if (op_in_sphere(400.f, B[b], t)) close = 1;              /* :442-474 */
                    }
                }
            body_alive[b] = (uint8_t)(!collide && close);
        }
        op_grid_free(&g);
        /* eliminateFarTarget :476-502 — against the surviving bodies */
        op_f3* live = (op_f3*)malloc((nb ? nb : 1) * sizeof(op_f3));
        size_t nl = 0;
        for (size_t b = 0; b < nb; b++) if (body_alive[b]) live[nl++] = B[b];
        op_grid gb;
        op_grid_build(&gb, live, nl, 100.f);
        for (size_t t = 0; t < nt; t++) {
            int cx0, cx1, cy0, cy1, keep = 0;
            op_grid_range(&gb, T[t].x, T[t].y, 401.f, &cx0, &cx1, &cy0, &cy1);
            for (int cy = cy0; cy <= cy1 && !keep; cy++)
                for (int cx = cx0; cx <= cx1 && !keep; cx++) {
                    size_t c = (size_t)cy * gb.nx + cx;
                    for (int s = gb.start[c]; s < gb.start[c + 1]; s++)
                        if (op_in_sphere(400.f, T[t], live[gb.idx[s]])) { keep = 1; break; }
                }
            if (keep) targ[ntk++] = T[t];
        }
        op_grid_free(&gb);
        free(live);
    } else {
        memcpy(targ, T, nt * sizeof(op_f3));
        ntk = nt;
    }

    int* todo = (int*)malloc(nb * sizeof(int));
    op_f3* body_rot = (op_f3*)malloc(nb * sizeof(op_f3));
    op_f3* targ_rot = (op_f3*)malloc((ntk ? ntk : 1) * sizeof(op_f3));
    op_leg_t* legs_rot = (op_leg_t*)malloc((size_t)nlegs * sizeof(op_leg_t));
    const float reach = op_max_reach(legs, nlegs);

    for (int o = 0; o < nq; o++) {
        size_t ntodo = 0;
        for (size_t b = 0; b < nb; b++)
            if (body_alive[b] && !standable[b]) todo[ntodo++] = (int)b; /* flipWorkingSide :396-399 */
        if (ntodo == 0) break;
        op_f4 q = {quats[4 * o], quats[4 * o + 1], quats[4 * o + 2], quats[4 * o + 3]};
        for (size_t k = 0; k < ntodo; k++) body_rot[todo[k]] = op_qt_rotate(q, B[todo[k]]); /* :401-411 */
        for (size_t t = 0; t < ntk; t++) targ_rot[t] = op_qt_rotate(q, targ[t]);
        for (int l = 0; l < nlegs; l++) legs_rot[l] = op_rotate_leg_data(q, legs[l]);       /* :743-760 */
        op_grid g;
        op_grid_build(&g, targ_rot, ntk, 128.f);

        /* cull cylinders from legsWorking[0] after the limit rotation, :505-520 */
        const op_leg_t d = legs_rot[0];
        float s_p = sinf(d.coxa_pitch), c_p = cosf(d.coxa_pitch);
        op_pose_job proto;
        memset(&proto, 0, sizeof proto);
        proto.radius_in = d.body + c_p * d.coxa_length + d.femur_length + d.tibia_length;
        float plus_abs = d.tibia_length * sinf(d.tibia_absolute_pos) +
                         d.femur_length * sinf(fminf(OP_PI / 2, d.max_angle_femur));
        proto.plus_in = s_p * d.coxa_length + plus_abs;
        proto.minus_in = s_p * d.coxa_length - d.femur_length - d.tibia_length;
        proto.radius_out = d.body;
        proto.body_rot = body_rot; proto.targ_rot = targ_rot; proto.grid = &g;
        proto.legs = legs_rot; proto.nlegs = nlegs; proto.q = q; proto.reach = reach;
        proto.todo = todo; proto.standable = standable; proto.mark = (uint8_t)(o + 1);

        pthread_t th[64];
        op_pose_job jobs[64];
        int use = (ntodo < 64) ? 1 : threads;
        size_t chunk = (ntodo + (size_t)use - 1) / (size_t)use;
        int started = 0;
        for (int t = 0; t < use; t++) {
            size_t lo = (size_t)t * chunk, hi = lo + chunk < ntodo ? lo + chunk : ntodo;
            if (lo >= hi) break;
            jobs[t] = proto; jobs[t].lo = lo; jobs[t].hi = hi;
            if (use == 1) { op_pose_worker(&jobs[t]); break; }
            pthread_create(&th[t], NULL, op_pose_worker, &jobs[t]);
            started++;
        }
        for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
        op_grid_free(&g);
    }
    free(todo); free(body_rot); free(targ_rot); free(legs_rot); free(targ); free(body_alive);
}

/* ------------------------------------------------------------------------------------------ */
/* Body-space octree (several_leg_octree.cu, octree_util.cu.h) — sequential restatement.       */
/* The reference builds this tree with dynamic parallelism, device malloc and unsynchronised   */
/* shared flags, so its GPU result is not deterministic; this single-threaded      */
/* restatement is the arbiter: flags are OR-ed over all work items of a pass and written after  */
/* the pass, children already valid BEFORE the pass are skipped (several_leg_octree.cu:57-60).  */

#define OP_MINBOX 100.0f       /* settings.h:17 */
#define OP_DEADQUADRAN 128     /* settings.h:30 */
#define OP_ROT_BELOW 50.0f     /* settings.h:32 */
#define OP_CONVEX_RADIUS 100.0f

typedef struct { op_f3 center, half; } op_box;

static unsigned op_reverse_bits(unsigned b) { /* octree_util.cu.h:78-84 */
    b = (((b & 0xaaaaaaaau) >> 1) | ((b & 0x55555555u) << 1));
    b = (((b & 0xccccccccu) >> 2) | ((b & 0x33333333u) << 2));
    b = (((b & 0xf0f0f0f0u) >> 4) | ((b & 0x0f0f0f0fu) << 4));
    b = (((b & 0xff00ff00u) >> 8) | ((b & 0x00ff00ffu) << 8));
    return (b >> 16) | (b << 16);
}
static unsigned op_shift_between(unsigned v, unsigned lo, unsigned up, unsigned shift) { /* :42-61 */
    unsigned mask = ((1u << (up + 1)) - 1) ^ ((1u << lo) - 1);
    unsigned in = v & mask;
    unsigned sh = (in << shift) & mask;
    return (v & ~mask) | sh;
}
/* octree_util.cu.h:105-151 with SUB_QUAD = 1 */
static unsigned op_child_box(op_box parent, op_box* child, unsigned quad_count, unsigned index,
                             const uint8_t* small, int* missing) {
    *child = parent;
    unsigned quadr = op_reverse_bits(index) >> (32 - quad_count);
    float* off[3] = {&child->half.x, &child->half.y, &child->half.z};
    unsigned sub = quadr & ((1u << quad_count) - 1 + (1u << quad_count));
    float div[3] = {2, 2, 2};
    unsigned upper = quad_count - 1;
    *missing = 0;
    for (unsigned q = 0; q < 3; q++) {
        if (*off[q] < OP_MINBOX) {
            (*missing)++;
            if (((sub >> upper) & 1) && !small[q]) {
                *missing = OP_DEADQUADRAN;
                return 0;
            }
            sub = op_shift_between(sub, q, 3, 1);
            if (small[q]) upper++;
            div[q] = 1;
        }
    }
    op_f3 old = child->half;
    child->half.x = child->half.x / div[0];
    child->half.y = child->half.y / div[1];
    child->half.z = child->half.z / div[2];
    op_f3 mv = {old.x - child->half.x, old.y - child->half.y, old.z - child->half.z};
    if (sub & 1) mv.x *= -1; /* flipVectorOnQuad :91-103 */
    if (sub & 2) mv.y *= -1;
    if (sub & 4) mv.z *= -1;
    child->center.x += mv.x; child->center.y += mv.y; child->center.z += mv.z;
    return sub;
}
unsigned op_create_child_box(const float* parent6, unsigned child, const uint8_t* small3,
                             float* child6, int* missing_quad) {
    op_box p, c;
    memcpy(&p, parent6, sizeof p);
    unsigned r = op_child_box(p, &c, 3, child, small3, missing_quad);
    memcpy(child6, &c, sizeof c);
    return r;
}
static int op_in_box(op_f3 v, op_f3 half) { /* octree_util.cu.h:153-159 */
    float ex = fabsf(half.x), ey = fabsf(half.y), ez = fabsf(half.z);
    return ex >= v.x && ey >= v.y && ez >= v.z && -ex < v.x && -ey < v.y && -ez < v.z;
}
int op_is_in_box(const float* v3, const float* box6) {
    op_f3 v = {v3[0], v3[1], v3[2]}, h = {box6[3], box6[4], box6[5]};
    return op_in_box(v, h);
}

typedef struct op_node {
    op_box box;
    int validity, leaf, raw, on_edge;
    struct op_node* children; /* 8 when allocated */
} op_node;

static int op_null_box(op_box b) {
    return b.center.x == 0 && b.center.y == 0 && b.center.z == 0 && b.half.x == 0 && b.half.y == 0 &&
           b.half.z == 0;
}

/* validity_child, several_leg_octree.cu:19-151 */
static void op_validity_pass(op_node* parent, const op_f3* foot, size_t nt, const op_leg_t* leg) {
    int on_edge[8] = {0}, valid_leaf[8] = {0}, valid[8] = {0};
    const int rot = parent->box.half.x < OP_ROT_BELOW;
    const float margin = rot ? 0.f : OP_ROT_BELOW / 3;
    const int n_angle = rot ? 27 : 1;
    const float reach = leg->body + leg->coxa_length + leg->femur_length + leg->tibia_length;
    const op_f3 elong = {parent->box.half.x + reach, parent->box.half.y + reach, parent->box.half.z + reach};
    for (int c = 0; c < 8; c++) {
        op_node* node = &parent->children[c];
        if (node->validity) continue; /* DEADQUADRAN or already processed */
        const op_box nb = node->box;
        const float edge_raw = nb.half.x * nb.half.x + nb.half.y * nb.half.y + nb.half.z * nb.half.z;
        for (size_t t = 0; t < nt; t++) {
            op_f3 vect = {foot[t].x - nb.center.x, foot[t].y - nb.center.y, foot[t].z - nb.center.z};
            if (!op_in_box(vect, elong)) continue;
            for (int a = 0; a < n_angle; a++) {
                float q4[4];
                op_quaternion_from_angle_index((unsigned)a, q4);
                op_f4 quat = {q4[0], q4[1], q4[2], q4[3]};
                int reach_count = 0, cross_count = 0;
                for (int k = 0; k < 4; k++) {
                    op_f3 v = vect;
                    op_leg_t l = *leg;
                    l.body_angle = OP_PI / 4 * k; /* LegMount, settings.h:41-42 */
                    int sub = op_distance_global(&v, &l, quat);
                    int cross;
                    if (edge_raw > OP_CONVEX_RADIUS * OP_CONVEX_RADIUS)
                        cross = op_in_box(v, nb.half); /* :99-103 (the margin box is built and dropped) */
                    else
                        cross = (v.x * v.x + v.y * v.y + v.z * v.z) < edge_raw + margin;
                    cross_count += cross;
                    reach_count += sub;
                }
                int edge = cross_count > 0;            /* LegCount - LegNumberForStab = 0 */
                int reachability = parent->validity || reach_count >= 4;
                if (edge) on_edge[c] = 1;
                if (reachability) valid[c] = 1;
                if (reachability && !edge) valid_leaf[c] = 1;
            }
        }
    }
    for (int c = 0; c < 8; c++) { /* :134-150 */
        op_node* node = &parent->children[c];
        if (valid[c]) node->validity = 1;
        if (valid_leaf[c]) node->leaf = 1;
        if (on_edge[c] && !valid_leaf[c]) node->on_edge = 1;
    }
}

/* branchKernel, several_leg_octree.cu:241-377 */
static void op_branch(op_node* parent, const op_f3* foot, size_t nt, const op_leg_t* leg) {
    if (parent->raw) {
        parent->children = (op_node*)calloc(8, sizeof(op_node));
        const uint8_t small[3] = {0, 0, 0};
        for (unsigned i = 0; i < 8; i++) {
            op_node* node = &parent->children[i];
            op_box nb;
            int missing;
            op_child_box(parent->box, &nb, 3, i, small, &missing);
            if (missing == OP_DEADQUADRAN) {
                node->leaf = 1; node->raw = 0; node->validity = 1; node->on_edge = 1;
                memset(&node->box, 0, sizeof node->box);
                continue;
            }
            node->on_edge = 0; node->validity = 0;
            node->box = nb;
            if (3 - missing <= 0) { node->leaf = 1; node->raw = 0; }
            else { node->leaf = 0; node->raw = 1; }
        }
        parent->raw = 0;
        op_validity_pass(parent, foot, nt, leg);
        return;
    }
    for (int i = 0; i < 8; i++) { /* go deeper :296-313 */
        op_node* node = &parent->children[i];
        if (!node->on_edge) node->leaf = 1;
        if (!node->leaf) op_branch(node, foot, nt, leg);
    }
}
/* One refinement pass on a hand-built parent: the children as branchKernel initialises them
 * (several_leg_octree.cu:315-352), then validity_child (:19-151).  flags: 8 x {validity, leaf, raw,
 * onEdge}; boxes: 8 x {center, topOffset} — the layout oracle/ref_gpu_shim.cu::refgpu_validity_child
 * reports for the reference's own kernel. */
void op_validity_children(const float* parent_box6, int parent_validity, const float* footholds, size_t nt,
                          const op_leg_t* leg, uint8_t* flags, float* boxes) {
    op_node parent;
    memset(&parent, 0, sizeof parent);
    memcpy(&parent.box, parent_box6, sizeof parent.box);
    parent.validity = parent_validity != 0;
    parent.raw = 1;
    op_branch(&parent, (const op_f3*)footholds, nt, leg);
    for (int c = 0; c < 8; c++) {
        const op_node* n = &parent.children[c];
        flags[4 * c + 0] = (uint8_t)n->validity, flags[4 * c + 1] = (uint8_t)n->leaf;
        flags[4 * c + 2] = (uint8_t)n->raw, flags[4 * c + 3] = (uint8_t)n->on_edge;
        memcpy(boxes + 6 * c, &n->box, sizeof n->box);
    }
    free(parent.children);
}

/* fill_recus, octree_util.cu:123-147 */
static size_t op_collect(const op_node* node, float* out, size_t n, size_t cap) {
    for (int i = 0; i < 8; i++) {
        const op_node* ch = &node->children[i];
        int endpoint = !(ch->leaf || ch->raw || op_null_box(ch->box));
        int is_valid = !op_null_box(ch->box) && (ch->leaf || ch->raw) && ch->validity;
        if (endpoint) n = op_collect(ch, out, n, cap);
        else if (is_valid) {
            if (n < cap) { out[3 * n] = ch->box.center.x; out[3 * n + 1] = ch->box.center.y; out[3 * n + 2] = ch->box.center.z; }
            n++;
        }
    }
    return n;
}
static void op_free_tree(op_node* node) {
    if (!node->children) return;
    for (int i = 0; i < 8; i++) op_free_tree(&node->children[i]);
    free(node->children);
}
/* apply_oct, several_leg_octree.cu:391-488, with the compile-time MAX_DEPTH as a parameter */
size_t op_apply_oct(const float* footholds, size_t nt, const op_leg_t* leg, int max_depth,
                    float* out_xyz, size_t cap) {
    op_node root;
    memset(&root, 0, sizeof root);
    root.box.half.x = root.box.half.y = root.box.half.z = 5000.f; /* settings.h:26 */
    root.raw = 1;
    for (int d = 0; d < max_depth; d++) op_branch(&root, (const op_f3*)footholds, nt, leg);
    size_t n = root.children ? op_collect(&root, out_xyz, 0, cap) : 0;
    op_free_tree(&root);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Single-leg distance-field octree painted on the query array: recursive_kernel + fillOutKernel */
/* (one_leg_global.cu:168-251, octree_util.cu:9-26) driven like apply_recurs                     */
/* (cross_compiled.cu:82-139), sequentially.  Output: (depth of the leaf box holding the point,  */
/* 0, 0); points outside the root box keep what the caller put in `out` (the reference leaves    */
/* uninitialised device memory there).                                                           */
static void op_recurs_node(op_box box, const op_f3* in, size_t n, const op_leg_t* leg, op_f4 quat,
                           op_f3* out, int depth, int max_depth) {
    uint8_t small[3] = {fabsf(box.half.x) < OP_MINBOX, fabsf(box.half.y) < OP_MINBOX,
                        fabsf(box.half.z) < OP_MINBOX};
    unsigned quad_count = 3 - small[0] - small[1] - small[2];
    unsigned n_child = 1u << quad_count;
    for (unsigned ci = 0; ci < n_child; ci++) {
        op_box nb;
        int missing;
        op_child_box(box, &nb, quad_count, ci, small, &missing);
        if (missing == OP_DEADQUADRAN) continue;
        int too_small = (3 - missing) <= 0;
        op_f3 d = nb.center;
        op_distance_global(&d, leg, quat);
        int edge_in_box = op_norm3(d) < op_norm3(nb.half);
        if (edge_in_box && !too_small && depth < max_depth) {
            op_recurs_node(nb, in, n, leg, quat, out, depth + 1, max_depth);
        } else {
            for (size_t i = 0; i < n; i++) { /* fillOutKernel */
                op_f3 delta = {in[i].x - nb.center.x, in[i].y - nb.center.y, in[i].z - nb.center.z};
                if (op_in_box(delta, nb.half)) { out[i].x = (float)depth; out[i].y = 0; out[i].z = 0; }
            }
        }
    }
}
void op_apply_recurs(const float* xyz, size_t n, const op_leg_t* leg, const float* quat4, int max_depth,
                     float* out_xyz) {
    op_box root;
    memset(&root, 0, sizeof root);
    root.half.x = root.half.y = root.half.z = 5000.f;
    op_f4 q = {quat4[0], quat4[1], quat4[2], quat4[3]};
    op_recurs_node(root, (const op_f3*)xyz, n, leg, q, (op_f3*)out_xyz, 0, max_depth);
}
