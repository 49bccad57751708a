// TEST INFRASTRUCTURE.  C entry points over the reference's vendored orocos_kdl, compiled from the sources where they
// lie under /root/reference (oracle/kdl_ref/Makefile) into oracle/_ref/libkdl_ik.so.  It builds the Panda arm chain the
// way ycb_render/robotPose/kdl_parser.py does from the URDF (one Segment per joint: Joint(origin, axis, RotAxis) +
// the parent->joint frame as tip) and runs the solvers robot_kinematics constructs
// (ycb_render/robotPose/robot_pykdl.py:140-146): ChainFkSolverPos_recursive, ChainIkSolverVel_pinv,
// ChainIkSolverPos_NR_JL with default maxiter / eps.
#include <cstring>

#include "chain.hpp"
#include "chainfksolverpos_recursive.hpp"
#include "chainiksolverpos_nr_jl.hpp"
#include "chainiksolvervel_pinv.hpp"
#include "frames.hpp"
#include "rigidbodyinertia.hpp"

namespace KDL {
// The two inertia classes only ride along inside Segment (never read on the IK path); their own translation units
// use Eigen expression templates, so the few symbols Segment's constructor links against are provided here instead.
RotationalInertia::RotationalInertia(double Ixx, double Iyy, double Izz, double Ixy, double Ixz, double Iyz) {
    data[0] = Ixx; data[1] = data[3] = Ixy; data[2] = data[6] = Ixz; data[4] = Iyy; data[5] = data[7] = Iyz; data[8] = Izz;
}
RotationalInertia::~RotationalInertia() {}
RigidBodyInertia::RigidBodyInertia(double m_, const Vector &h_, const RotationalInertia &I_, bool) : m(m_), h(h_), I(I_) {}
RigidBodyInertia::RigidBodyInertia(double m_, const Vector &c_, const RotationalInertia &Ic) : m(m_), h(m * c_), I(Ic) {}
}  // namespace KDL

using namespace KDL;

struct KdlIk {
    Chain chain;
    ChainFkSolverPos_recursive *fk;
    ChainIkSolverVel_pinv *vel;
    ChainIkSolverPos_NR_JL *pos;
};

// frames: [num_segments][16] row-major 4x4 parent->joint transforms (URDF joint origins); axes: [num_segments][3] in the
// joint frame; movable[num_segments]: 1 = revolute, 0 = fixed.
extern "C" void *kdl_ik_create(const double *frames, const double *axes, const int *movable, int num_segments,
                               const double *q_min, const double *q_max) {
    KdlIk *k = new KdlIk();
    int nj = 0;
    for (int s = 0; s < num_segments; ++s) {
        const double *m = frames + 16 * s;
        Frame f(Rotation(m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]), Vector(m[3], m[7], m[11]));
        if (movable[s]) {
            Vector ax(axes[3 * s], axes[3 * s + 1], axes[3 * s + 2]);
            k->chain.addSegment(Segment(Joint(f.p, f.M * ax, Joint::RotAxis), f));   // kdl_parser.py: urdf_joint_to_kdl_joint
            ++nj;
        } else {
            k->chain.addSegment(Segment(Joint(Joint::None), f));
        }
    }
    JntArray lo(nj), hi(nj);
    for (int j = 0; j < nj; ++j) { lo(j) = q_min[j]; hi(j) = q_max[j]; }
    k->fk = new ChainFkSolverPos_recursive(k->chain);
    k->vel = new ChainIkSolverVel_pinv(k->chain);
    k->pos = new ChainIkSolverPos_NR_JL(k->chain, lo, hi, *k->fk, *k->vel);
    return k;
}

extern "C" void kdl_ik_destroy(void *h) {
    KdlIk *k = (KdlIk *)h;
    delete k->pos; delete k->vel; delete k->fk; delete k;
}

// robot_kinematics.inverse_kinematics (robot_pykdl.py:257-289): position [3], orientation quaternion xyzw [4], seed [nj]
// -> result [nj]; returns KDL's status (>= 0: solution found).
extern "C" int kdl_ik_solve(void *h, const double *position, const double *quat_xyzw, const double *seed, double *result) {
    KdlIk *k = (KdlIk *)h;
    const int nj = (int)k->chain.getNrOfJoints();
    JntArray q0(nj), q(nj);
    for (int j = 0; j < nj; ++j) q0(j) = seed[j];
    Frame goal(Rotation::Quaternion(quat_xyzw[0], quat_xyzw[1], quat_xyzw[2], quat_xyzw[3]),
               Vector(position[0], position[1], position[2]));
    const int rc = k->pos->CartToJnt(q0, goal, q);
    for (int j = 0; j < nj; ++j) result[j] = q(j);
    return rc;
}

// forward kinematics of the chain tip: out [16] row-major 4x4
extern "C" int kdl_fk(void *h, const double *q_in, double *out) {
    KdlIk *k = (KdlIk *)h;
    const int nj = (int)k->chain.getNrOfJoints();
    JntArray q(nj);
    for (int j = 0; j < nj; ++j) q(j) = q_in[j];
    Frame f;
    const int rc = k->fk->JntToCart(q, f);
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out[4 * r + c] = f.M(r, c); out[4 * r + 3] = f.p(r); }
    out[12] = out[13] = out[14] = 0.0; out[15] = 1.0;
    return rc;
}
