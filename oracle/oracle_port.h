/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference hot path (see oracle_port.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (liblrm_b200.so) never links or calls it.
 */
#ifndef LRM_ORACLE_PORT_H
#define LRM_ORACLE_PORT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 14-float order as the reference's LegDimensions (HeaderCPP.h:19-52). */
typedef struct {
    float body_angle, body, coxa_pitch, coxa_length, tibia_length, femur_length;
    float tibia_absolute_pos, tibia_absolute_neg;
    float max_angle_coxa, min_angle_coxa, max_angle_tibia, min_angle_tibia;
    float max_angle_femur, min_angle_femur;
} op_leg_t;

typedef struct { float x, y, z; } op_f3;
typedef struct { float x, y, z, w; } op_f4;

/* static_variables.cpp:6-93; robot 0 = moonbot, 1 = M2 */
void op_get_leg(int robot, float azimuth, op_leg_t* out);

/* one_leg.cu:280-341 */
int op_reachability_circles(op_f3 p, const op_leg_t* leg);
int op_distance_circles(op_f3* p_inout, const op_leg_t* leg);
/* one_leg_global.cu:74-130 (host branch) */
int op_reachability_global(op_f3 p, const op_leg_t* leg, op_f4 quat);
int op_distance_global(op_f3* p_inout, const op_leg_t* leg, op_f4 quat);

/* slab loops (one_leg_global.cu:132-147), optionally over `threads` host threads */
void op_reach(const float* xyz, size_t n, const op_leg_t* leg, const float* quat4, uint8_t* out,
              int threads);
void op_dist(const float* xyz, size_t n, const op_leg_t* leg, const float* quat4, float* out_xyz,
             uint8_t* out_flag, int threads);

/* planar pieces, for table pinning */
int op_find_region(float x, float y, const op_leg_t* leg);
int op_insert_circles(float x, float y, const op_leg_t* leg, float* out16);
int op_insert_intersec(const op_leg_t* leg, float* out20);

/* quaternion helpers in the reference's (mixed) layouts, unified_math_cuda.cu.h:13-83 */
op_f3 op_qt_rotate(op_f4 q, op_f3 v);
op_f4 op_qt_invert(op_f4 q);
op_f4 op_qt_multiply(op_f4 a, op_f4 b);
op_f4 op_quat_from_vect_angle(op_f3 axis, float angle);
op_f3 op_rpy_from_quat(op_f4 q);
op_f4 op_rpy_to_quat(float r, float p, float y);
op_leg_t op_rotate_leg_data(op_f4 quat, op_leg_t leg);
/* pointer-style twins for ctypes */
void op_qt_rotate_p(const float* q4, const float* v3, float* out3);
void op_qt_multiply_p(const float* a4, const float* b4, float* out4);
void op_quat_from_vect_angle_p(const float* axis3, float angle, float* out4);
void op_rpy_to_quat_p(float r, float p, float y, float* out4);
void op_rotate_leg_data_p(const float* q4, const op_leg_t* leg, op_leg_t* out);

/* ---- multi-leg positionability (several_leg.cu:48-67, 326-877) ------------------------ */
/* The 45 orientation quaternions of robot_full_struct (several_leg.cu:811-857), in loop order. */
int op_full_struct_orientations(float* out_quat4 /* 45*4 */);
/* 27 orientation samples of the octree path (octree_util.cu.h:184-198). */
void op_quaternion_from_angle_index(unsigned idx, float* out4);

/* Standability of every body position, restating multi_rot_estimator:
 *   standable[b] = 1 + index of the first orientation that succeeds, 0 if none (or culled).
 * bodies/targets: N x 3 AoS float.  legs: nlegs x op_leg_t.  quats: nq x 4.
 * pre_cull != 0 applies the constructor culls (60 mm always-colliding sphere, 400 mm far body,
 * 400 mm far target; several_leg.cu:371-374,413-502).
 */
void op_standability(const float* bodies, size_t nb, const float* targets, size_t nt,
                     const op_leg_t* legs, int nlegs, const float* quats, int nq, int pre_cull,
                     uint8_t* standable, int threads);

/* ---- body-space octree (several_leg_octree.cu:19-151,241-377; octree_util.cu.h:105-159) */
unsigned op_create_child_box(const float* parent6, unsigned child, const uint8_t* small3,
                             float* child6, int* missing_quad);
int op_is_in_box(const float* v3, const float* box6);
/* Sequential restatement of apply_oct with `max_depth` levels (reference ships MAX_DEPTH 1).
 * Returns the number of valid leaves; centres (x,y,z) are written to out_xyz (capacity cap). */
size_t op_apply_oct(const float* footholds, size_t nt, const op_leg_t* leg, int max_depth,
                    float* out_xyz, size_t cap);

/* One pass of validity_child (several_leg_octree.cu:19-151) over the 8 children of one parent box
 * (children initialised as in branchKernel :315-352).  flags: 8 x {validity, leaf, raw, onEdge},
 * boxes: 8 x {center xyz, topOffset xyz}. */
void op_validity_children(const float* parent_box6, int parent_validity, const float* footholds, size_t nt,
                          const op_leg_t* leg, uint8_t* flags, float* boxes);

/* apply_recurs (cross_compiled.cu:82-139): octree of the single-leg distance field painted on the
 * query points, out = (leaf depth, 0, 0); points outside the +-5000 mm root box are left untouched. */
void op_apply_recurs(const float* xyz, size_t n, const op_leg_t* leg, const float* quat4, int max_depth,
                     float* out_xyz);

#ifdef __cplusplus
}
#endif
#endif
