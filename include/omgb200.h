/*
 * omgb200.h -- C ABI of libomgb200.so: the B200-native (sm_100a) CHOMP hot path of OMG-Planner.
 *
 * Plain C: device/host pointers, sizes, a cudaStream_t passed as void*.  No torch types.  Every entry
 * point returns 0 on success or a negative omgb_status; omgb_last_error() gives the message.  Nothing
 * here ever calls exit() (the reference does: layers/sdf_matching_loss_kernel.cu:241-246).
 * The caller owns all buffers; the library allocates device memory only inside omgb_scene_* calls.
 *
 * Reference interfaces replaced (file:line under the reference tree):
 *   omgb_sdf_loss              omg_cuda.sdf_loss_forward            layers/omg_layers.cpp:24-49,
 *                                                                   layers/sdf_matching_loss_kernel.cu:204-262
 *   omgb_scene_set_sdf         Env.combine_sdfs output              omg/core.py:366-411
 *   omgb_scene_set_objects     per-object parameters                omg/cost.py:303-335
 *   omgb_scene_set_robot       robot_kinematics constants + Robot   ycb_render/robotPose/robot_pykdl.py:98-112,
 *                                                                   omg/core.py:152-164
 *   omgb_scene_set_metric      cfg.Ainv + goal-set projection       omg/config.py:199-220, omg/optimizer.py:102-107
 *   omgb_chomp_step            Optimizer.optimize(force_update)     omg/optimizer.py:115-135 -> omg/cost.py:451-532
 *   omgb_chomp_plan            the fixed-goal inner loop of plan()  omg/planner.py:612-627
 *   omgb_chomp_step_host       same as omgb_chomp_step, HOST buffers (H2D + D2H inside the call)
 *   omgb_batch_obstacle_cost   Cost.batch_obstacle_cost             omg/cost.py:192-286
 *   omgb_goal_costs            Learner.cost_vector, device half     omg/online_learner.py:104-150 (omg/util.py:261-290,
 *                                                                   omg/cost.py:192-286 with arc_length, the two sums)
 *   omgb_chomp_plan_history    plan() with history_trajectories     omg/planner.py:605-628
 *   omgb_chomp_plan_step       one iteration of plan() when the     omg/planner.py:612-628
 *                              goal changes between iterations
 *   omgb_learner_update        Learner.update_goal after the        omg/online_learner.py:151-160, 162-249
 *                              collision costs (cost vector, FTL /
 *                              FTC / Exp / MD / Proj, goal selection)
 *   omgb_chomp_plan_goalset    plan() in goal-set mode with the     omg/planner.py:612-635
 *                              learner, the whole loop in one call
 *   omgb_traj_interpolate      Trajectory.interpolate_waypoints     omg/core.py:59-78 -> omg/util.py:238-258
 *   omgb_sdf_pack              SignedDensityField.from_pth/.resize  omg/sdf_tools.py:187-193, 37-39,
 *                              + Env.combine_sdfs                   omg/core.py:366-411
 *   omgb_point_sdf             PointEnv.compute_sdf_from_points     omg/core.py:426-457 (scipy cKDTree.query)
 *   omgb_ik_solve              solve_one_pose_ik for every grasp x  omg/planner.py:16-87, 296-455 ->
 *                              seed (robot.inverse_kinematics)      ycb_render/robotPose/robot_pykdl.py:257-289 ->
 *                                                                   orocos_kdl/src/chainiksolverpos_nr_jl.cpp:61-101
 *   omgb_hand_poses            forward_kinematics_parallel()[:, 7]  omg/planner.py:262-283 (hand frame only)
 */
#ifndef OMGB200_H_
#define OMGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMGB_VERSION 100
#define OMGB_NUM_LINKS 10      /* 7 arm links, hand, 2 fingers (omg/core.py:171-182) */
#define OMGB_NUM_DOF 9         /* 7 arm + 2 finger joints (omg/core.py:28) */
#define OMGB_MAX_OBJECTS 64
#define OMGB_MAX_BODY_POINTS 32
#define OMGB_INFO_STRIDE 16

typedef enum {
    OMGB_OK = 0,
    OMGB_ERR_INVALID = -1,   /* bad argument (what AT_ASSERT raised in layers/omg_layers.cpp:5-7) */
    OMGB_ERR_CUDA = -2,      /* CUDA runtime error (the reference exit(-1)s) */
    OMGB_ERR_STATE = -3,     /* scene not fully configured */
    OMGB_ERR_UNSUPPORTED = -4
} omgb_status;

/* Layout of the per-trajectory info row written by omgb_chomp_step ([B, OMGB_INFO_STRIDE] doubles);
 * the fields of the reference's info dict (omg/cost.py:509-530, omg/optimizer.py:173-174). */
enum {
    OMGB_INFO_OBS = 0,            /* info["obs"] (x n inflated in top-k mode, SURVEY A-3) */
    OMGB_INFO_SMOOTH = 1,         /* info["smooth"] */
    OMGB_INFO_COST = 2,           /* info["cost"] */
    OMGB_INFO_COLLIDE = 3,        /* info["collide"] (a count, SURVEY A-17) */
    OMGB_INFO_REACH = 4,          /* info["reach"] goal distance */
    OMGB_INFO_GRAD_NORM = 5,      /* info["grad"] */
    OMGB_INFO_WOBS_GRAD_NORM = 6, /* info["weighted_obs_grad"] */
    OMGB_INFO_WSMOOTH_GRAD_NORM = 7,
    OMGB_INFO_TERMINATE = 8,      /* after check_joint_limit (omg/optimizer.py:174) */
    OMGB_INFO_VIOLATE_LIMIT = 9,
    OMGB_INFO_EXECUTE = 10,
    OMGB_INFO_FAILURE_TERMINATE = 11,
    OMGB_INFO_P_IN = 12,          /* in-bounds (body point, enabled object) pairs this iteration */
    OMGB_INFO_NONZERO = 13,       /* body points with non-zero potential */
    OMGB_INFO_LIMIT_ROUNDS = 14,  /* rounds of handle_joint_limit taken */
    OMGB_INFO_RESERVED = 15
};

typedef struct omgb_scene omgb_scene_t;

/* Scalars Cost/Optimizer read from cfg on every call (omg/config.py:29-104); the host mirror snapshots
 * them per call because Optimizer.update() writes the schedules back into cfg (omg/optimizer.py:68-80). */
typedef struct {
    int32_t n_waypoints;           /* cfg.timesteps */
    int32_t goal_set_proj;         /* cfg.goal_set_proj */
    int32_t constraint_rows;       /* c: reach_tail_length with use_standoff, else 1; 0 when !goal_set_proj */
    int32_t top_k_collision;       /* cfg.top_k_collision; 0 = sum over all points */
    int32_t uncheck_finger_collision; /* cfg.uncheck_finger_collision (0 or -1) */
    int32_t consider_finger;       /* cfg.consider_finger: finger links in the top-k sum, finger DOFs updated */
    int32_t allow_collision_point; /* cfg.allow_collision_point */
    int32_t pre_terminate;         /* cfg.pre_terminate */
    int32_t joint_limit_max_steps; /* cfg.joint_limit_max_steps */
    int32_t update;                /* 0: info_only=True; 1: always update (force_update=True);
                                      2: update unless this iteration reports terminate (optimizer.py:124) */
    double time_interval;          /* cfg.time_interval */
    double obstacle_weight;        /* cfg.obstacle_weight (after Optimizer.update) */
    double smoothness_weight;      /* cfg.smoothness_weight */
    double step_size;              /* cfg.step_size */
    double clip_grad_scale;        /* cfg.clip_grad_scale */
    double terminate_smooth_loss;  /* cfg.terminate_smooth_loss */
    double link_smooth_weight[OMGB_NUM_DOF]; /* cfg.link_smooth_weight */
} omgb_step_params_t;

int omgb_version(void);
const char *omgb_last_error(void);

/* ---- scene: everything that is constant across CHOMP iterations, resident in HBM ------------------- */
int omgb_scene_create(omgb_scene_t **out, int device);
int omgb_scene_destroy(omgb_scene_t *scene);

/* Kinematic constants (HOST pointers, fp64): pose_0 [10,4,4], tip2joint [10,4,4], joint_axis [10,3],
 * center_offset [10,4,4] (robot_p3.pkl via robot_pykdl.py:101-110; "origin" is aliased to joint_axis as
 * in :104 unless use_true_joint_origin != 0, then joint_origin [10,3] must be given), body_points
 * [10,p,3] in the centre-offset link frames (Robot.collision_points, omg/core.py:166-190), padded joint
 * limits lower/upper [9] (omg/core.py:157-164). */
int omgb_scene_set_robot(omgb_scene_t *scene, const double *pose_0, const double *tip2joint,
                         const double *joint_axis, const double *joint_origin, int use_true_joint_origin,
                         const double *center_offset, const double *body_points, int points_per_link,
                         const double *lower, const double *upper);

/* Packed SDFs exactly as Env.combine_sdfs leaves them: d_sdf_grids DEVICE [O,X,Y,Z] fp32 (borrowed, not
 * copied: zero-copy on env.sdf_torch), h_sdf_limits HOST [O,10] fp32.  The grid preprocessing (lower-bound grid)
 * runs on `stream` -- the stream the grid was produced on -- and only that stream is waited for. */
int omgb_scene_set_sdf(omgb_scene_t *scene, const float *d_sdf_grids, const float *h_sdf_limits,
                       int num_objects, int dim_x, int dim_y, int dim_z, void *stream);

/* Optional data layout for the exact-evaluation path of the fused kernels (the layout of
 * omg/core.py:366-411 being replaced).  layout 0: read the reference [O,X,Y,Z] tensor as it is (default, zero-copy).
 * layout 1: additionally keep a bricked copy -- 8^3-cell bricks of 8 KB, one float4 per cell holding the cell's four
 * (y, z) taps -- so that a trilinear sample is two aligned 128-bit loads instead of eight scalar taps.  Same tap
 * values and interpolation order: results are bit-identical.  Costs 4x the grid bytes; call after omgb_scene_set_sdf
 * (which drops a stale copy) and before omgb_scene_set_objects.  Environment OMGB_SDF_LAYOUT=1 selects it by default. */
int omgb_scene_set_sdf_layout(omgb_scene_t *scene, int layout, void *stream);

/* Per-object parameters built by Cost.compute_obstacle_cost_layer (omg/cost.py:303-335), HOST arrays:
 * pose_inv [O,4,4] fp32 (world->object), epsilons, padding_scales, clearances, disables [O] fp32. */
int omgb_scene_set_objects(omgb_scene_t *scene, const float *pose_inv, const float *epsilons,
                           const float *padding_scales, const float *clearances, const float *disables,
                           void *stream);

/* Smoothness metric: h_Ainv HOST [n,n] fp64 (cfg.Ainv), h_proj HOST [n,c] fp64 = Ainv C^T (C Ainv C^T)^-1
 * (omg/optimizer.py:107) or NULL when c == 0. */
int omgb_scene_set_metric(omgb_scene_t *scene, int n_waypoints, const double *h_Ainv, int constraint_rows,
                          const double *h_proj);

/* Exact accelerations that never change results (tests compare on/off bit for bit): the lower-bound grid
 * culling and the longest-first CTA order.  -1 keeps the current value; both default to on. */
int omgb_scene_set_options(omgb_scene_t *scene, int use_lower_bound, int use_longest_first);

/* How omgb_chomp_step_host moves its HOST buffers: 0 (default) = zero-copy when every buffer lies in mapped pinned
 * memory (cudaHostAlloc / cudaHostRegister: the fused kernel reads its inputs and writes its results over PCIe
 * itself), otherwise staged copies pipelined over chunks of trajectories; 1 = staged, one piece; 2 = staged,
 * pipelined; 3 = zero-copy or OMGB_ERR_INVALID.  Results are identical in every mode. */
int omgb_scene_set_host_mode(omgb_scene_t *scene, int mode);

/* Number of kernels this library has launched so far in this process (bench.py reports the delta). */
unsigned long long omgb_launch_count(void);

/* Diagnostic: when d_phase_clocks (DEVICE int64 [B,16]) is non-NULL every CTA of the fused step writes clock64()
 * at its phase boundaries (profiles/ phase breakdowns); NULL (default) disables it. */
int omgb_scene_set_profile(omgb_scene_t *scene, long long *d_phase_clocks);

/* ---- the raw operator (drop-in for omg_cuda.sdf_loss_forward); all pointers DEVICE, fp32 ------------ */
size_t omgb_sdf_loss_workspace_bytes(int num_objects);
int omgb_sdf_loss(const float *pose_init, const float *sdf_grids, const float *sdf_limits,
                  const float *points, const float *epsilons, const float *padding_scales,
                  const float *clearances, const float *disables, int num_points, int num_objects,
                  int dim_x, int dim_y, int dim_z, float *potentials, float *potential_grads,
                  float *collides, void *workspace, void *stream);

/* ---- one fused CHOMP iteration over a batch of trajectories; all pointers DEVICE --------------------
 * xi [B,n,9] fp64 in/out (Trajectory.data), start [B,9], end [B,9] (traj.start / traj.end),
 * goal_rows [B,c,9] (reach_grasps[goal_idx] or goal_set[goal_idx]; NULL when c == 0),
 * active [B] uint8 or NULL (0 = leave this trajectory untouched),
 * grad_out [B,n,9] fp64 or NULL (info["gradient"]), info [B,OMGB_INFO_STRIDE] fp64,
 * dbg_potentials [B,n,10,p] fp32 or NULL, dbg_points [B,n,10,p,3] fp32 or NULL (collision_pts columns),
 * row_obs [B,n] fp64 or NULL, zero-initialised by the caller (obstacle part of info["cost_traj"]). */
int omgb_chomp_step(omgb_scene_t *scene, const omgb_step_params_t *params, int batch, double *xi,
                    const double *start, const double *end, const double *goal_rows, const uint8_t *active,
                    double *grad_out, double *info, float *dbg_potentials, float *dbg_points, double *row_obs,
                    void *stream);

/* `iters` iterations with per-iteration schedules (HOST arrays [iters]: obstacle_weight, smoothness_weight,
 * step_size as Optimizer.update() would set them).  stop_on_terminate != 0 freezes a trajectory once an
 * iteration with index > 0 reports terminate (omg/planner.py:627); done [B] uint8 DEVICE scratch/out. */
int omgb_chomp_plan(omgb_scene_t *scene, const omgb_step_params_t *params, int iters,
                    const double *obstacle_weights, const double *smoothness_weights, const double *step_sizes,
                    int stop_on_terminate, int batch, double *xi, const double *start, const double *end,
                    const double *goal_rows, uint8_t *done, double *info, void *stream);

/* omgb_chomp_plan that also records what Planner.plan keeps per iteration (omg/planner.py:605-628):
 * hist_xi DEVICE [iters,B,n,9] fp64 = xi after every iteration (history_trajectories[1:]; the caller holds
 * history_trajectories[0], the initial xi), hist_info DEVICE [iters,B,OMGB_INFO_STRIDE] = every iteration's info row.
 * Either may be NULL.  A trajectory frozen by stop_on_terminate repeats its last row in later slots. */
int omgb_chomp_plan_history(omgb_scene_t *scene, const omgb_step_params_t *params, int iters,
                            const double *obstacle_weights, const double *smoothness_weights,
                            const double *step_sizes, int stop_on_terminate, int batch, double *xi,
                            const double *start, const double *end, const double *goal_rows, uint8_t *done,
                            double *info, double *hist_xi, double *hist_info, void *stream);

/* Same as omgb_chomp_step with HOST buffers (pinned or pageable); xi and info are valid on return (the stream is
 * synchronised).  Transfer strategy: omgb_scene_set_host_mode. */
int omgb_chomp_step_host(omgb_scene_t *scene, const omgb_step_params_t *params, int batch, double *h_xi,
                         const double *h_start, const double *h_end, const double *h_goal_rows,
                         double *h_info, void *stream);

/* Cost.batch_obstacle_cost (omg/cost.py:192-286): joints DEVICE [M,9] fp64 (rad);
 * arc_length <= 0: plain potentials; > 0: joints are G groups of arc_length configurations, potentials
 * are multiplied by the workspace speed |x_i - x_{i-1}|/dt with x_{-1} = FK(start [9]) (omg/config.py:162-187);
 * out potentials [M,10,p], grads [M,10,p,3] (or NULL), collides [M,10,p] fp32 DEVICE. */
int omgb_batch_obstacle_cost(omgb_scene_t *scene, const double *joints, int num_configs, int arc_length,
                             const double *start, double time_interval, int uncheck_finger_collision,
                             float *potentials, float *grads, float *collides, void *stream);

/* Goal scoring for the online learner (omg/online_learner.py:104-150), fused: for every trajectory b and goal g,
 *   costs[b,g] = sum over i < arc_length, links, body points of  potential(x_i) * |x_i - x_{i-1}| / dt
 * where the configurations q_i, i = 0..arc_length-1, are the interior points of the joint-space line from
 * from[b] (= traj.data[start], q_{-1}) to goals[b,g] (multi_interpolate_waypoints, mode "linear") and x are the
 * body points under forward kinematics.  from: DEVICE rows of 9 fp64, row b at from + b*from_stride (pass
 * xi + start*9 with from_stride = n*9 to score from waypoint `start` of every trajectory in xi [B,n,9]);
 * goals: DEVICE [B,G,9] fp64, or [G,9] shared by all trajectories when goals_shared != 0; costs: DEVICE [B,G] fp32.
 * The Learner passes uncheck_finger_collision = 0 (online_learner.py:138). */
int omgb_goal_costs(omgb_scene_t *scene, int batch, const double *from, long long from_stride, const double *goals,
                    int num_goals, int goals_shared, int arc_length, double time_interval,
                    int uncheck_finger_collision, float *costs, void *stream);

/* One iteration (index `iteration`, 0-based) of a plan whose goal rows may change between iterations (goal-set mode
 * with the online learner): omgb_chomp_step with the plan's bookkeeping -- trajectories with done[b] != 0 are frozen,
 * done[b] is set when iteration > 0 reports terminate (stop_on_terminate), hist_xi / hist_info (layouts of
 * omgb_chomp_plan_history, may be NULL) receive slot `iteration`.  The schedule comes from params. */
int omgb_chomp_plan_step(omgb_scene_t *scene, const omgb_step_params_t *params, int iteration,
                         int stop_on_terminate, int batch, double *xi, const double *start, const double *end,
                         const double *goal_rows, uint8_t *done, double *info, double *hist_xi, double *hist_info,
                         void *stream);

/* ---- the online learner's goal re-weighting (omg/online_learner.py:151-249), all pointers DEVICE ------------------ */
enum {
    OMGB_LEARNER_FTL = 0, OMGB_LEARNER_FTC = 1, OMGB_LEARNER_EXP = 2, OMGB_LEARNER_MD = 3, OMGB_LEARNER_PROJ = 4,
    OMGB_LEARNER_INIT = 5   /* Learner.__init__ (:91-102): goal = argmin of the cost vector, no state change */
};

typedef struct {
    int32_t alg;                  /* cfg.ol_alg as one of the enum values above */
    int32_t num_goals;            /* G = len(traj.goal_set), <= 256 */
    int32_t n_waypoints;          /* cfg.timesteps */
    int32_t first_waypoint;       /* the waypoint the goal lines start from (:108-110) */
    int32_t constraint_rows;      /* c of the goal rows written out (>= 1) */
    int32_t normalize_cost;       /* cfg.normalize_cost */
    double base_obstacle_weight;  /* cfg.base_obstacle_weight */
    double smoothness_base_weight;/* cfg.smoothness_base_weight */
    double dist_eps;              /* cfg.dist_eps */
    double eta;                   /* Learner.eta = sqrt(log(G + 1) / optim_steps) (Exp) */
    double etas[5];               /* Learner.etas (MD experts) */
} omgb_learner_params_t;

/* For every trajectory b: cost vector from collision [B,G] fp32 (omgb_goal_costs) and the joint-difference term
 * between xi[b, first_waypoint] and goal_set[b, g]; update of the goal distribution; goal_idx[b] = argmax p;
 * end[b] = goal_set[b, goal_idx]; goal_rows[b] = reach[b, goal_idx] ([c,9]) or, when reach is NULL, goal_set[b, goal_idx]
 * repeated c times (c = 1 without standoff).
 * goal_set: [B,G,9] or [G,9] when goals_shared; reach: [B,G,c,9] / [G,c,9] or NULL.
 * State (in/out, the Learner's fields): p [B,G], sum_costs [B,G], experts_p [B,5,G], experts_costs [B,5], q [B,5], fp64.
 * done [B] uint8 or NULL: trajectories whose plan has ended keep their goal.  cost_vector [B,G] fp64 out or NULL.
 * selected [B] int32 out or NULL (the iteration's slot of Planner.selected_goals). */
int omgb_learner_update(const omgb_learner_params_t *params, int batch, const double *xi, const float *collision,
                        const double *goal_set, int goals_shared, const double *reach, double *p, double *sum_costs,
                        double *experts_p, double *experts_costs, double *q, const uint8_t *done, int32_t *goal_idx,
                        double *end, double *goal_rows, double *cost_vector, int32_t *selected, void *stream);

/* Device buffers of a goal-set plan with goal switching (all DEVICE; the arguments omgb_goal_costs, omgb_learner_update
 * and omgb_chomp_plan_step take, with the same layouts). */
typedef struct {
    double *xi;                 /* [B,n,9] in/out */
    const double *start;        /* [B,9] */
    double *end;                /* [B,9] in/out: the selected goal (omgb_learner_update) */
    double *goal_rows;          /* [B,c,9] in/out: the selected goal's projection rows */
    uint8_t *done;              /* [B] in/out: trajectories whose plan has ended (zeroed by the caller) */
    double *info;               /* [B,16] out: the last iteration's info rows */
    double *hist_xi;            /* [iters,B,n,9] or NULL */
    double *hist_info;          /* [iters,B,16] or NULL */
    const double *goal_set;     /* [B,G,9], or [G,9] when goals_shared */
    const double *reach;        /* [B,G,c,9] / [G,c,9] or NULL (no standoff) */
    const double *reach_goals;  /* [B,G,9] / [G,9]: the configurations the goal lines end at (reach[..., -1, :] with
                                   standoff, else goal_set; omg/online_learner.py:121-125) */
    double *p, *sum_costs, *experts_p, *experts_costs, *q;   /* the Learner's state (omgb_learner_update) */
    int32_t *goal_idx;          /* [B] in/out */
    int32_t *selected;          /* [learner_iters,B] out or NULL: Planner.selected_goals */
    float *collision;           /* [B,G] fp32 scratch (the goal costs of the current iteration) */
    int32_t goals_shared;
    int32_t reserved_;
} omgb_goalset_plan_buffers_t;

/* Planner.plan's loop in goal-set mode with the online learner (omg/planner.py:612-635), enqueued by one call: for
 * t = 0 .. iters-1:  if t < learner_iters (cfg.optim_steps): omgb_goal_costs from waypoint first_waypoint[t] (skipped
 * for OMGB_LEARNER_PROJ) and omgb_learner_update (params `learner` with that first_waypoint);  then
 * omgb_chomp_plan_step(iteration t, stop_on_terminate) with obstacle_weight / smoothness_weight / step_size =
 * schedule[t] (HOST [iters,3], Optimizer.update's values for that iteration, omg/optimizer.py:59-80).
 * Results are those of the three calls made one at a time; nothing synchronises unless timeout_s >= 0 (cfg.timeout,
 * planner.py:629): then the host clock is compared every 8 iterations against the device's progress and the loop
 * stops early.  *iters_enqueued (HOST, may be NULL) = iterations enqueued. */
int omgb_chomp_plan_goalset(omgb_scene_t *scene, const omgb_step_params_t *params,
                            const omgb_learner_params_t *learner, int iters, int learner_iters,
                            const double *schedule, const int32_t *first_waypoint, int batch,
                            const omgb_goalset_plan_buffers_t *buffers, double timeout_s, int *iters_enqueued,
                            void *stream);

/* ---- trajectory initialisation and the SDF asset path (no scene; they run on the calling thread's current
 * CUDA device) ------------------------------------------------------------------------------------------------ */

/* omg/util.py:238-258 for a batch: waypoints DEVICE [B,K,9] fp64 at knots linspace(0,1,K) (the reference always
 * passes K = 2: start and end), sampled at the n interior points of linspace(0,1,n+2) -> xi DEVICE [B,n,9].
 * mode 1 = scipy CubicSpline(bc_type="clamped") (cfg.traj_interpolate = "cubic"), mode 0 = interp1d "linear". */
int omgb_traj_interpolate(const double *waypoints, int batch, int num_knots, int n_waypoints, int mode,
                          double *xi, void *stream);

/* One object's raw signed-distance grid as loaded from disk (DEVICE pointer). */
typedef struct {
    const void *data;
    int32_t shape[3];   /* logical [X,Y,Z] (SignedDensityField.data.shape) */
    int32_t layout;     /* 0: stored [X,Y,Z]; 1: stored [Y,X,Z] -- the .pth files hold sdf_torch[0,0] =
                           data.permute(1,0,2) (real_world/convert_sdf.py:43, undone by omg/sdf_tools.py:191) */
    int32_t dtype;      /* 0: fp32, 1: fp64 */
    float scale;        /* SignedDensityField.resize ratio (omg/sdf_tools.py:37-39, fp32 multiply); 1 = none */
} omgb_sdf_source_t;

/* Env.combine_sdfs (omg/core.py:366-411): sources HOST [O] -> d_out DEVICE [O,X,Y,Z] fp32, every grid in the corner
 * of its slot, the rest 1.0.  (The [O,10] limits are host arithmetic on 10 numbers per object; the host mirror
 * computes them with the reference's own expression.) */
int omgb_sdf_pack(const omgb_sdf_source_t *sources, int num_objects, int dim_x, int dim_y, int dim_z, float *d_out,
                  void *stream);

/* PointEnv.compute_sdf_from_points (omg/core.py:426-457): for every voxel (gx[i], gy[j], gz[k]) the distance to the
 * nearest of d_points [N,3] (all DEVICE fp64) -> d_out32 [X,Y,Z] fp32 and/or d_out64 fp64 (either may be NULL). */
int omgb_point_sdf(const double *d_points, int num_points, const double *d_gx, const double *d_gy,
                   const double *d_gz, int dim_x, int dim_y, int dim_z, float *d_out32, double *d_out64, void *stream);

/* ---- goal-set construction: batched inverse kinematics ---------------------------------------------
 * The reference solves IK with KDL's ChainIkSolverPos_NR_JL (Newton-Raphson with joint limits, <= 100 steps, 1e-6 twist
 * tolerance) over ChainIkSolverVel_pinv (truncated pseudo-inverse from SVD_HH) for the chain panda_link0 -> panda_hand
 * (robot_pykdl.py:114-146), one grasp pose and one seed at a time in a 4-process pool (omg/planner.py:400-436).
 * omgb_ik_solve runs the same iteration for every (pose, seed) pair in one launch.
 *
 * chain_frames HOST [8,4,4] fp64: parent->joint transforms of the 7 arm joints and the fixed hand joint
 *   (robot_kinematics._pose_0[:8]); every joint turns about its local z (the URDF's axis="0 0 1").
 * q_min / q_max HOST [7]: the padded joint limits handed to the solver (robot_pykdl.py:123-138).
 * d_targets DEVICE [P, T, 7] fp64: for pose p the chain of T targets (position xyz, quaternion xyzw) solved in order,
 *   each seeded with the previous solution (solve_one_pose_ik: the last standoff pose first, then the reach tail);
 *   T = 1 for a single solve.
 * d_seeds DEVICE [S, 7].  Outputs: d_sols DEVICE [P, S, T, 7]; d_solved DEVICE int32 [P, S] = number of solves that
 *   succeeded before the first failure (== T: the whole chain solved); d_steps DEVICE int32 [P, S, T] or NULL =
 *   Newton steps of each solve attempted.
 * Two builds of the same arithmetic exist (factorisation in registers / in local memory); the library picks by
 * problem size, the environment variable OMGB_IK_SVD=reg|local forces one.  Results are identical. */
int omgb_ik_solve(const double *chain_frames, const double *q_min, const double *q_max, const double *d_targets,
                  int num_poses, int chain_length, const double *d_seeds, int num_seeds, double *d_sols,
                  int *d_solved, int *d_steps, void *stream);

/* Hand frame (panda_hand in the base frame) of num_configs joint vectors: d_joints DEVICE fp64, row m at
 * d_joints + m * joint_stride (first 7 entries used) -> d_poses DEVICE [M,4,4] row-major. */
int omgb_hand_poses(const double *chain_frames, const double *d_joints, long long joint_stride, int num_configs,
                    double *d_poses, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OMGB200_H_ */
