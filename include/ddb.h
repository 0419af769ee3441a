/*
 * ddb.h — C ABI of the B200-native batched differentiable timestep.
 *
 * This is the drop-in boundary for the path behind the reference's
 *   dart::neural::forwardPass            (dart/neural/NeuralUtils.hpp:45-46, NeuralUtils.cpp:26-66)
 *   dart::neural::BackpropSnapshot       (dart/neural/BackpropSnapshot.hpp:22-803)
 *   simulation::World::step              (dart/simulation/World.cpp:206-279)
 * executed for N independent worlds that share one skeleton topology.
 *
 * Conventions
 *  - All arithmetic is FP64, index/flag data is int32.
 *  - Every pointer named *_dev is a CUDA device pointer on the handle's device;
 *    the *_host entry points take ordinary host pointers and do the copies
 *    themselves (that is the "e2e" path a binding with host buffers calls).
 *  - Batched vectors use the layout chosen in ddb_params.layout:
 *      DDB_LAYOUT_WORLD_MAJOR  x[world * n_dofs + dof]   (a [N, n] row-major tensor)
 *      DDB_LAYOUT_DOF_MAJOR    x[dof * n_worlds + world] (structure of arrays)
 *    Batched matrices are [world][col][row] (column-major per world, as Eigen
 *    returns them from BackpropSnapshot::get*Jacobian) in both layouts.
 *  - Every function returns 0 on success, a DDB_ERR_* code otherwise;
 *    ddb_last_error() returns a human-readable host string.
 *  - There is no CPU fallback: without a CUDA device ddb_create fails.
 *  - Threading: a handle is not thread-safe (neither is a World); different handles may be driven from different
 *    host threads, also on one device: entry points that launch kernels serialise per device while they enqueue work.
 */
#ifndef DDB_H_
#define DDB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDB_ABI_VERSION 1
#define DDB_MAX_BODIES 48
#define DDB_MAX_DOFS 48

enum {
  DDB_OK = 0,
  DDB_ERR_INVALID = 1,     /* bad argument / unsupported topology */
  DDB_ERR_CUDA = 2,        /* a CUDA runtime call failed           */
  DDB_ERR_NO_DEVICE = 3,   /* no CUDA device: there is no CPU path */
  DDB_ERR_STATE = 4        /* e.g. backward before forward         */
};

/* Joint kinds on the hot path (dart/dynamics/{Weld,Revolute,Prismatic,Free}Joint.cpp). */
enum {
  DDB_JOINT_WELD = 0,      /* 0 DOF */
  DDB_JOINT_REVOLUTE = 1,  /* 1 DOF, RevoluteJoint.cpp:140-189 */
  DDB_JOINT_PRISMATIC = 2, /* 1 DOF, PrismaticJoint.cpp:139-187 */
  DDB_JOINT_FREE = 3       /* 6 DOF, FreeJoint.cpp:64-81,587-601,673-693 */
};

/* Collision shapes handled by the in-tree narrow phase (dart/collision/dart/DARTCollide.cpp). */
enum { DDB_SHAPE_NONE = 0, DDB_SHAPE_BOX = 1, DDB_SHAPE_SPHERE = 2,
       DDB_SHAPE_CAPSULE = 3 /* axis = local z; against boxes only, and only where an end cap is the closest feature (see ddb_collide_pairs) */ };

enum { DDB_LAYOUT_WORLD_MAJOR = 0, DDB_LAYOUT_DOF_MAJOR = 1 };

/* A rigid transform: rotation row-major r[0..8], then translation p[0..2]. */
typedef struct ddb_transform {
  double r[9];
  double p[3];
} ddb_transform;

/*
 * One BodyNode together with its parent Joint. A superset of the reference's own
 * flat-array descriptor JointAndBody (dart/dynamics/SimpleFeatherstone.hpp:16-30).
 * Bodies are listed parent-before-child; all skeletons of the world are
 * concatenated (parent = -1 marks a skeleton root), DOFs are numbered in that
 * order exactly as World::getPositions() concatenates them (World.cpp:1027-1040).
 * Springs and damping do not exist here: World::addSkeleton zeroes them
 * (World.cpp:698-732).
 */
typedef struct ddb_body {
  int32_t parent;      /* index of the parent body, -1 for a root            */
  int32_t joint_type;  /* DDB_JOINT_*                                        */
  int32_t dof_offset;  /* first generalized coordinate of the parent joint   */
  int32_t skeleton;    /* skeleton id (constraint grouping, mobility)        */
  int32_t gravity_mode;/* BodyNode gravity mode (BodyNode.cpp:1694)          */
  int32_t shape_type;  /* DDB_SHAPE_*; collidable iff != NONE                */
  double axis[3];      /* joint axis in the joint frame (1-DOF joints)       */
  ddb_transform parent_to_joint; /* Joint::mT_ParentBodyToJoint              */
  ddb_transform child_to_joint;  /* Joint::mT_ChildBodyToJoint               */
  double mass;
  double com[3];       /* Inertia::mCenterOfMass                             */
  double moment[6];    /* Ixx Iyy Izz Ixy Ixz Iyz (Inertia.cpp:394-409)      */
  double friction;     /* BodyNodeAspect.hpp:47 default 1.0                  */
  double restitution;  /* BodyNodeAspect.hpp:48 default 0.0                  */
  double shape_size[3];/* box: full extents; sphere: radius in [0]; capsule: radius in [0], height of the cylinder part in [1] */
  ddb_transform shape_offset;    /* ShapeNode relative transform             */
} ddb_body;

typedef struct ddb_topology {
  int32_t n_bodies;
  int32_t n_dofs;
  const ddb_body* bodies;
  /* per-DOF limits used only by clipLossGradientsToBounds
     (BackpropSnapshot.cpp:374-428); NULL means +-infinity */
  const double* position_lower;
  const double* position_upper;
  const double* velocity_lower;
  const double* velocity_upper;
  const double* force_lower;
  const double* force_upper;
} ddb_topology;

/* World-level parameters; ddb_default_params() fills the reference defaults
   (World.cpp:67-94, ContactConstraint.cpp:43-55). */
typedef struct ddb_params {
  double gravity[3];              /* (0,0,-9.81)  World.cpp:71 */
  double time_step;               /* 1e-3         World.cpp:72 */
  double contact_clipping_depth;  /* 0.03         World.cpp:84 */
  int32_t layout;                 /* DDB_LAYOUT_*                */
  int32_t n_snapshot_slots;       /* >=1; slots retained for T-step rollouts */
  int32_t reserved[6];
} ddb_params;

typedef struct ddb_handle ddb_handle;

void ddb_default_params(ddb_params* p);
int ddb_abi_version(void);

/* Replaces: World construction + World::addSkeleton for N clones (World.cpp:109-157,683-760). */
int ddb_create(const ddb_topology* topo, const ddb_params* params, int32_t n_worlds,
               int32_t device, ddb_handle** out);
int ddb_destroy(ddb_handle* h);
const char* ddb_last_error(const ddb_handle* h);
int32_t ddb_num_worlds(const ddb_handle* h);
int32_t ddb_num_dofs(const ddb_handle* h);

/*
 * Replaces neural::forwardPass(world, idempotent=false) for every world
 * (NeuralUtils.cpp:26-66 -> World::step World.cpp:206-279). Inputs are the
 * pre-step positions, velocities and joint forces; outputs the post-step state.
 * The snapshot needed by ddb_backward / ddb_jacobians is retained in the
 * handle's current slot (see ddb_snapshot_select).
 * lcp_cache_inout / n_rows_inout / active_set_out / status_out may be NULL for
 * contact-free topologies.
 */
int ddb_forward(ddb_handle* h, void* cuda_stream,
                const double* q_dev, const double* qd_dev, const double* tau_dev,
                double* q_next_dev, double* qd_next_dev,
                double* lcp_cache_inout_dev, int32_t* n_rows_inout_dev,
                int32_t* active_set_out_dev, int32_t* status_out_dev);

/*
 * Replaces BackpropSnapshot::backprop (BackpropSnapshot.cpp:115-186) including
 * clipLossGradientsToBounds (:374-428): given dL/dq', dL/dqd' returns
 * dL/dq, dL/dqd, dL/dtau of the step retained in the current slot.
 */
int ddb_backward(ddb_handle* h, void* cuda_stream,
                 const double* g_q_next_dev, const double* g_qd_next_dev,
                 double* g_q_dev, double* g_qd_dev, double* g_tau_dev);

/*
 * Replaces BackpropSnapshot::get{PosPos,PosVel,VelPos,VelVel,ForceVel}Jacobian
 * (BackpropSnapshot.cpp:432-735,1237-1360). Each output is [N][n][n]
 * column-major per world; any pointer may be NULL to skip that matrix.
 */
int ddb_jacobians(ddb_handle* h, void* cuda_stream,
                  double* pos_pos_dev, double* pos_vel_dev, double* vel_pos_dev,
                  double* vel_vel_dev, double* force_vel_dev);

/*
 * Contact stage (run inside ddb_forward when the topology has collidable pairs): collision detection, contact
 * constraints, boxed-LCP assembly by impulse tests, reduce -> Dantzig -> validity -> PGS -> friction-free PGS, impulses
 * (ConstraintSolver::solve, constraint/ConstraintSolver.cpp:386-404; BoxedLcpConstraintSolver.cpp:188-684 without the
 * gradient-matrix branches when ddb_params.reserved[1] == 1; by default the gradient-enabled solve of forwardPass:
 * LCPUtils::guessSolution on a size change, short-circuit through the cached solution, classification into
 * CLAMPING / UPPER_BOUND / NOT_CLAMPING / ILLEGAL and least-squares standardisation of the clamping impulses,
 * ConstrainedGroupGradientMatrices.cpp:167-324,467-697). ddb_params.reserved[0] caps the contacts per world
 * (default 64); rows = 3 * contacts.
 * status_out of ddb_forward = solver tier mask per world (bit 0 short-circuit, 1 Dantzig, 2 PGS, 3 friction dropped,
 * 8 capacity hit, 9 a capsule pair in a configuration this path does not carry: its contacts are missing from the step); active_set_out [N][max_rows] = ConstraintMapping per LCP row (-1 CLAMPING, -2 NOT_CLAMPING,
 * -3 ILLEGAL, -4 IRRELEVANT = no such row, >= 0 UPPER_BOUND pointing at its normal row).
 */
int32_t ddb_contact_capacity(const ddb_handle* h, int32_t* max_contacts, int32_t* max_rows);
/*
 * Replaces World::getCachedLCPSolution / setCachedLCPSolution (BoxedLcpConstraintSolver.cpp:174-186; saved and restored by
 * RestorableSnapshot.cpp:23,32 for forwardPass(world, idempotent=true), NeuralUtils.cpp:29-33,61-62) for every world: the
 * solver's cached solution the handle carries from step to step when ddb_forward is given no cache of the caller's.
 * cache [N][max_rows] doubles, n_rows [N] (device pointers, copies ordered on `stream`).
 */
int ddb_lcp_cache_get(ddb_handle* h, void* cuda_stream, double* cache_out_dev, int32_t* n_rows_out_dev);
int ddb_lcp_cache_set(ddb_handle* h, void* cuda_stream, const double* cache_dev, const int32_t* n_rows_dev);
int ddb_contact_debug_host(ddb_handle* h, int32_t* counts, int32_t world, double* contacts, int32_t* bodies, int32_t* dims,
                           double* A, double* b, double* lo, double* hi, double* x, int32_t* findex);

/* With DDB_CONTACT_PROFILE=1 in the environment: clock ticks per section of the contact stage summed over worlds since the
   last call (host array [16]: collision+constraints, assembly, guessSolution, short-circuit, Dantzig, reduce for PGS,
   post-PGS, commit). DDB_ERR_STATE when profiling is off. */
int ddb_contact_profile_host(int64_t* ticks_out);

/* Statistics of the last ddb_forward: Gauss-Seidel sweeps the PGS fallback spent per world (host array [N]). */
int ddb_contact_pgs_sweeps_host(ddb_handle* h, int32_t* sweeps_out);
/* ... and the size (rows of the reduced problem) of the largest problem each world handed to it; 0 = none (host array [N]). */
int ddb_contact_pgs_rows_host(ddb_handle* h, int32_t* rows_out);

/* Replaces BackpropSnapshot::getInvMassMatrix / getMassMatrix helpers
   (python/_diffdart/neural/BackpropSnapshot.cpp:42-146): [N][n][n] col-major. */
int ddb_inv_mass_matrix(ddb_handle* h, void* cuda_stream, double* minv_dev);

/* Snapshot slots for T-step rollouts (the shape of SingleShot::getSnapshots,
   dart/trajectory/SingleShot.cpp:719-760 and its reverse sweep :650-700). */
int ddb_snapshot_select(ddb_handle* h, int32_t slot);
int32_t ddb_snapshot_current(const ddb_handle* h);
/* grow the slot pool to at least n_slots (never shrinks); returns the pool size in *n_out */
int ddb_snapshot_reserve(ddb_handle* h, int32_t n_slots, int32_t* n_out);

/*
 * T-step rollouts: the caller of the hot path in the reference's trajectory code.
 *   ddb_rollout_forward   SingleShot::getSnapshots (dart/trajectory/SingleShot.cpp:719-760): T steps from (q0, qd0) under
 *                         tau[t]; step t is retained in snapshot slot t (the pool grows as needed) and its post-step state
 *                         is written to q_seq[t], qd_seq[t].
 *   ddb_rollout_backward  SingleShot::backpropGradientWrt (SingleShot.cpp:650-700): the reverse sweep over the retained steps.
 *                         g_q_seq[t], g_qd_seq[t] = loss gradient w.r.t. the state after step t (zeros where the loss does not
 *                         look); the gradient pulled back by the later steps is added before step t is pulled back. Outputs:
 *                         dL/dq0, dL/dqd0 and dL/dtau[t] for every step.
 * All sequences are [T] consecutive batched vectors in the handle's layout. Contact-free handles run the rollout as chunks
 * of worlds on the handle's lanes (ddb_set_lanes, default 2), with no synchronisation between chunks until the call's end;
 * handles with a contact stage step the whole batch in order on `stream`.
 */
int ddb_rollout_forward(ddb_handle* h, void* stream, int32_t T, const double* q0, const double* qd0,
                        const double* tau, double* q_seq, double* qd_seq);
int ddb_rollout_backward(ddb_handle* h, void* stream, int32_t T, const double* g_q_seq,
                         const double* g_qd_seq, double* g_q0, double* g_qd0, double* g_tau);
/* number of internal streams (1..16) used by the rollout entry points and the host-buffer path */
int ddb_set_lanes(ddb_handle* h, int32_t n_lanes);

/* Host-buffer entry points: same semantics, host pointers, copies inside (contact-free handles with world-major vectors
   pipeline copy-in / kernel / copy-out over chunks of worlds on the lanes). */
int ddb_forward_host(ddb_handle* h, const double* q, const double* qd, const double* tau,
                     double* q_next, double* qd_next);
int ddb_backward_host(ddb_handle* h, const double* g_q_next, const double* g_qd_next,
                      double* g_q, double* g_qd, double* g_tau);
int ddb_jacobians_host(ddb_handle* h, double* pos_pos, double* pos_vel, double* vel_pos,
                       double* vel_vel, double* force_vel);

/*
 * Batched boxed LCP  A x = b + w  (lo <= x <= hi, friction rows bounded by hi*x[findex]):
 * replaces BoxedLcpSolver::solve(n, A, x, b, nub, lo, hi, findex, earlyTermination)
 * (dart/constraint/BoxedLcpSolver.hpp:92-101) for n_problems independent problems, as
 * BoxedLcpConstraintSolver calls it (nub = 0; BoxedLcpConstraintSolver.cpp:425-486,499-546).
 *   DDB_LCP_DANTZIG  DantzigBoxedLcpSolver -> ODE dSolveLCP (dart/external/odelcpsolver/lcp.cpp:780-1114),
 *                    same pivots and the same floating-point operation order: results are bit-identical.
 *   DDB_LCP_PGS      PgsBoxedLcpSolver::solve (PgsBoxedLcpSolver.cpp:79-262), options (10000, 1e-10, 1e-8, 1e-8);
 *                    x_inout is the warm start.
 * A is [n_problems][n_max][n_max] row-major (only the leading n_rows[i] x n_rows[i] block is read; n_rows may be
 * NULL = n_max everywhere); b, lo, hi, x are [n_problems][n_max]; ok_out[i] = the solver's bool (may be NULL).
 */
enum { DDB_LCP_DANTZIG = 0, DDB_LCP_PGS = 1 };
int ddb_lcp_solve(int32_t device, void* cuda_stream, int32_t n_problems, int32_t n_max, const int32_t* n_rows_dev,
                  const double* A_dev, const double* b_dev, const double* lo_dev, const double* hi_dev,
                  const int32_t* findex_dev, int32_t solver, int32_t early_termination,
                  double* x_inout_dev, int32_t* ok_out_dev);
/*
 * Batched Lemke solver for standard LCPs  w = M z + q >= 0, z >= 0, z.w = 0 : replaces
 * int dart::lcpsolver::Lemke(const MatrixXd& M, const VectorXd& q, VectorXd* z) (dart/lcpsolver/Lemke.hpp:44-46, Lemke.cpp:61-337)
 * for n_problems independent problems (complementary pivoting with the artificial variable, the reference's tolerances 1e-5 / 1e-8,
 * 1000 pivots at most, final validate() with threshold 1e-4). M is [n_problems][n_max][n_max] row-major (only the leading
 * n_rows[i] x n_rows[i] block is read; n_rows may be NULL = n_max everywhere, n_max <= 128); q, z_out are [n_problems][n_max];
 * err_out[i] = the reference's return code (0 solved, 1 iteration limit, 2 ray termination, 3 validation failed, 4 no pivot; z = 0
 * unless 0 or 3); pivots_out[i] = pivots taken (may be NULL). Not on the step path (only ODELCPSolver calls it).
 */
int ddb_lemke_solve(int32_t device, void* cuda_stream, int32_t n_problems, int32_t n_max, const int32_t* n_rows_dev,
                    const double* M_dev, const double* q_dev, double* z_out_dev, int32_t* err_out_dev, int32_t* pivots_out_dev);

/* LCPUtils::isLCPSolutionValid (dart/constraint/LCPUtils.cpp:12-81, tolerance 1e-5) per problem. */
int ddb_lcp_valid(int32_t device, void* cuda_stream, int32_t n_problems, int32_t n_max, const int32_t* n_rows_dev,
                  const double* A_dev, const double* x_dev, const double* b_dev, const double* lo_dev, const double* hi_dev,
                  const int32_t* findex_dev, int32_t ignore_friction_indices, int32_t* valid_out_dev);

/*
 * Narrow phase for n_pairs independent shape pairs: replaces dart::collision::collide(o1, o2, result)
 * (dart/collision/dart/DARTCollide.cpp:4675-5183) for box / sphere shapes — collideBoxBox -> dBoxBox (:713-1427, up to
 * 8 contacts per pair), collideBoxSphere / collideSphereBox (:1429-1733), collideSphereSphere (:1735-1800).
 * shape: DDB_SHAPE_*; size [n][3] (box extents / radius in [0]); T [n][12] (R row-major, p).
 * contacts_out [n][8][26]: point, normal, edgeAClosest, edgeAFixed, edgeADir, edgeBClosest, edgeBFixed, edgeBDir (3 each),
 * penetrationDepth, ContactType (Contact.hpp:43-80); counts_out [n]. Bit-identical to the reference's box-box.
 * Capsules (DDB_SHAPE_CAPSULE, size = radius, height): collideCapsuleBox / collideBoxCapsule (:4217-4361) hand the pair to the sphere-box
 * routine with the sphere of the capsule end that the contact lies beyond; the reference finds the contact position with libccd's MPR
 * (not part of the reference tree), here the end is the one from which the distance of the capsule's axis to the box grows strictly.
 * Where the closest feature is the cylinder part (the reference's pipe contacts, createCapsuleMeshContact :2887-3300) and for
 * capsule-capsule / capsule-sphere pairs the pair is NOT carried: counts_out = -1, and a step reports bit 9 of its status.
 */
int ddb_collide_pairs(int32_t device, void* cuda_stream, int32_t n_pairs, const int32_t* shape1_dev, const double* size1_dev,
                      const double* T1_dev, const int32_t* shape2_dev, const double* size2_dev, const double* T2_dev,
                      double* contacts_out_dev, int32_t* counts_out_dev);

/* FP64 FMA micro-benchmark (8 independent DFMA chains per thread, every SM busy): the measured
   FP64-pipe peak of this GPU in TFLOP/s, the denominator of the FP64 roofline fraction. */
int ddb_measure_fp64_peak(int32_t device, double* tflops_out);

/* Counts kernels launched by this handle since creation (bench "gpu_launches"). */
int64_t ddb_launch_count(const ddb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* DDB_H_ */
