/*
 * mpinets_b200.h -- C ABI of libmpinets_b200.so (B200 / sm_100a rollout engine for Motion Policy Networks).
 *
 * The reference (NVlabs/motion-policy-networks) has no FFI: its boundary for this path is a set of Python
 * callables.  Each entry point below replaces the arithmetic behind one of them; the
 * Python shim in mpinets_b200/ re-exports the reference's names over this ABI (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every function returns 0 on success, a negative mpn_status otherwise; mpn_last_error() gives the
 *     thread-local message of the last failure;
 *   - all tensor arguments are caller-owned DEVICE pointers (fp32 / int32 / u8), dense row-major, sizes explicit;
 *     the library never frees or reallocates caller memory; it owns only its context workspace;
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises the host except
 *     mpn_ctx_create/destroy, mpn_set_robot_tables, mpn_load_weight, mpn_weights_finalize and mpn_reserve;
 *   - one context per device; a context is not thread-safe.
 */
#ifndef MPINETS_B200_H
#define MPINETS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpn_ctx mpn_ctx;

typedef enum {
  MPN_OK = 0,
  MPN_ERR_INVALID = -1,   /* bad argument (shape, null pointer, unsupported size) */
  MPN_ERR_CUDA = -2,      /* a CUDA runtime call failed */
  MPN_ERR_STATE = -3,     /* tables / weights not loaded yet */
  MPN_ERR_NOMEM = -4
} mpn_status;

/* arithmetic used by the set-abstraction MLPs / FC head */
typedef enum {
  MPN_PREC_FP32 = 0,      /* fp32 SIMT FMA, fp32 accumulate: the 1e-5 parity mode */
  MPN_PREC_BF16 = 1,      /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM: the throughput mode */
  MPN_PREC_BF16X3 = 2     /* split-bf16 operands (x = hi + lo; a*w = a_hi w_hi + a_lo w_hi + a_hi w_lo) on tcgen05, fp32 accumulate:
                             the parity-grade tensor-core mode -- delta-q within 1e-5 of the fp32 reference (model.py:75-91 in fp32) */
} mpn_precision;

typedef struct {
  int32_t n_robot;        /* run_inference.py:52  NUM_ROBOT_POINTS    (2048) */
  int32_t n_obstacle;     /* run_inference.py:53  NUM_OBSTACLE_POINTS (4096) */
  int32_t n_target;       /* run_inference.py:54  NUM_TARGET_POINTS   (128)  */
  int32_t max_cuboids;    /* M1: padded cuboid rows per problem   (data_loader.py:198-215) */
  int32_t max_cylinders;  /* M2: padded cylinder rows per problem */
  int32_t quirk_frames;   /* 1: replicate geometry.py:213/:444 row-2 quirk (default), 0: textbook rotation */
  uint64_t seed;          /* counter-based RNG seed (Philox4x32-10 key) */
} mpn_config;

const char* mpn_last_error(void);
const char* mpn_version(void);

int mpn_ctx_create(int device, const mpn_config* cfg, mpn_ctx** out);
int mpn_ctx_destroy(mpn_ctx* ctx);
/* pre-size the context workspace for `max_batch` problems (otherwise grown lazily; growing is illegal during
 * CUDA-graph capture) */
int mpn_reserve(mpn_ctx* ctx, int max_batch);

/* robofin tables as data (HOST pointers, copied): FrankaRealRobot.JOINT_LIMITS [7][2] (utils.py:50,84,192);
 * FrankaSampler canonical link points [P][3] + link id [P] (link0..7, hand, leftfinger, rightfinger = 0..10);
 * gripper points in the right_gripper frame [Pe][3] (sample_end_effector); FrankaCollisionSampler spheres. */
int mpn_set_robot_tables(mpn_ctx* ctx, const float* joint_limits, int n_link_points, const float* link_points,
                         const int32_t* link_ids, int n_ee_points, const float* ee_points, int n_spheres,
                         const float* sphere_centers, const float* sphere_radii, const int32_t* sphere_links,
                         float prismatic);

/* weights by reference state-dict key (model.py:47-66,385-393; e.g. "point_cloud_encoder.SA_modules.0.mlps.0.0.weight"),
 * HOST fp32 pointer, copied.  mpn_weights_finalize() checks completeness and builds the packed device copies. */
int mpn_load_weight(mpn_ctx* ctx, const char* name, const float* host_data, const int64_t* shape, int ndim);
int mpn_weights_finalize(mpn_ctx* ctx);

/* ---- pointnet2_ops._ext replacements (model.py:27,365-383) -------------------------------- */
/* furthest_point_sampling: xyz [B][N][stride>=3] -> idx [B][npoint]; new_xyz [B][npoint][3] optional (gather fused) */
int mpn_fps(mpn_ctx* ctx, void* stream, const float* xyz, int B, int N, int stride, int npoint, int32_t* idx,
            float* new_xyz);
/* ball_query: first `nsample` indices (in index order) with d^2 < r^2, padded with the first hit */
int mpn_ball_query(mpn_ctx* ctx, void* stream, float radius, int nsample, const float* xyz, int B, int N, int stride,
                   const float* new_xyz, int npoint, int32_t* idx);
/* gather_points: feat [B][C][N], idx [B][m] -> out [B][C][m] */
int mpn_gather_points(mpn_ctx* ctx, void* stream, const float* feat, int B, int C, int N, const int32_t* idx, int m,
                      float* out);
/* group_points: feat [B][C][N], idx [B][m][ns] -> out [B][C][m][ns] */
int mpn_group_points(mpn_ctx* ctx, void* stream, const float* feat, int B, int C, int N, const int32_t* idx, int m,
                     int ns, float* out);
/* PointnetSAModule.forward fused (FPS -> ball query -> group -> shared MLP -> max): module = 0,1,2 (model.py:365-383).
 * xyz [B][N][stride], feats point-major [B][N][C_in]; outputs new_xyz [B][npoint][3] (NULL for module 2),
 * new_feats point-major [B][npoint][C_out].  fps_idx / ball_idx optional debug outputs (may be NULL). */
int mpn_sa_forward(mpn_ctx* ctx, void* stream, int module, int precision, const float* xyz, int stride,
                   const float* feats, int feat_stride, int B, int N, float* new_xyz, float* new_feats,
                   int32_t* fps_idx, int32_t* ball_idx);

/* ---- robofin replacements ----------------------------------------------------------------------------------- */
/* FrankaSampler FK: q [B][7] -> link frames [B][11][12] (3x4 row-major), right_gripper pose [B][12] (either may be NULL) */
int mpn_fk(mpn_ctx* ctx, void* stream, const float* q, int B, float* frames, float* eef);
/* FrankaSampler.sample(q, n): writes rows [0,n) of cloud [B][rows][4] as (x,y,z,0); subset keyed by (seed, step) */
int mpn_sample_robot(mpn_ctx* ctx, void* stream, const float* q, int B, int n, uint32_t step, float* cloud, int rows);
/* FrankaSampler.sample_end_effector(poses, n, frame="right_gripper") (run_inference.py:113-116, data_loader.py:158-161,
 * planning_node.py:71-74): n gripper-surface points (hand + fingers, table given in the right_gripper frame) transformed by each
 * pose [B][12] (3x4 row-major) -> out [B][n][3].  The subset is keyed by (seed, problem0 + b) exactly like the target rows of
 * mpn_build_cloud, i.e. out[b] == cloud[b][n_robot + n_obstacle : ][:, :3] for n = n_target. */
int mpn_sample_end_effector(mpn_ctx* ctx, void* stream, const float* poses, int B, int n, uint32_t problem0, float* out);
/* FrankaCollisionSampler.compute_spheres: q [B][7] -> world centres [B][S][3] */
int mpn_compute_spheres(mpn_ctx* ctx, void* stream, const float* q, int B, float* centers);
/* utils.(un)normalize_franka_joints with the loaded limits */
int mpn_normalize_joints(mpn_ctx* ctx, void* stream, const float* q, int n, float* q_norm);
int mpn_unnormalize_joints(mpn_ctx* ctx, void* stream, const float* q_norm, int n, float* q);

/* ---- mpinets.geometry replacements -------------------------------------------------------------------------- */
typedef struct {
  const float* cuboid_centers;   /* [B][M1][3] */
  const float* cuboid_dims;      /* [B][M1][3] */
  const float* cuboid_quats;     /* [B][M1][4] wxyz */
  const float* cylinder_centers; /* [B][M2][3] */
  const float* cylinder_radii;   /* [B][M2] */
  const float* cylinder_heights; /* [B][M2] */
  const float* cylinder_quats;   /* [B][M2][4] */
} mpn_scene;

/* TorchCuboids/TorchCylinders.sdf (geometry.py:238-288,456-507): points [B][N][3] -> sdf [B][N];
 * which = 0 min(both), 1 cuboids only, 2 cylinders only; all-masked -> +inf */
int mpn_sdf_points(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, const float* points, int N, int which,
                   float* sdf);
/* construct_mixed_point_cloud + run_inference.make_point_cloud_from_primitives (geometry.py:571-608,
 * run_inference.py:93-134): q0 [B][7] unnormalised, target [B][12] right_gripper pose -> cloud [B][Nr+No+Nt][4] */
int mpn_build_cloud(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, const float* q0, const float* target,
                    uint32_t problem0, float* cloud);
/* mpn_build_cloud with an explicit RNG counter per problem (problem_ids u32 [B], device) instead of problem0 + b: dataset batches
 * (PointCloudBase.get_inputs, data_loader.py:237-278) key each item by its dataset index, so a sample's cloud does not depend on the
 * batch it lands in, the worker or the rank. */
int mpn_build_cloud_ids(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, const float* q0, const float* target,
                        const uint32_t* problem_ids, float* cloud);
/* Training-time joint noise of PointCloudBase.get_inputs (data_loader.py:167-180): q_out = clamp(q + random_scale * N(0, 1),
 * FrankaRealRobot.JOINT_LIMITS), q_norm_out = normalize(q_out); q, q_out, q_norm_out [B][7].  The normals are Box-Muller transforms
 * of Philox4x32-10(counter = (sample id, epoch, 6, pair), key = seed); sample_ids u32 [B] (device; NULL: 0..B-1). */
int mpn_augment_joints(mpn_ctx* ctx, void* stream, const float* q, int B, float random_scale, const uint32_t* sample_ids,
                       uint32_t epoch, float* q_out, float* q_norm_out);
/* Planner.clean_point_cloud (interactive_demo/mpinets_ros/nodes/planning_node.py:187-228): keep the points of xyz [N][3] inside the
 * task-tabletop or mount-table box, then a random subset without replacement of n_out of them (keyed by (seed, cloud_id)) ->
 * out_xyz [n_out][3] (and out_rgba [n_out][4] from rgba [N][4], both optional).  kept[0] (device int) receives the number of points
 * inside the workspace; when it is < n_out nothing is written (np.random.choice raises there).  scratch: int32 [N] (device). */
int mpn_clean_point_cloud(mpn_ctx* ctx, void* stream, const float* xyz, const float* rgba, int N, int n_out, uint32_t cloud_id,
                          float* out_xyz, float* out_rgba, int32_t* kept, int32_t* scratch);
/* run_inference.make_point_cloud_from_problem (run_inference.py:58-90), for problems that carry an obstacle point cloud
 * (PlanningProblem.obstacle_point_cloud, mpinets_types.py:44): obstacle rows = a random subset WITHOUT replacement of
 * obstacle_points [B][max_points][3] restricted to the first obstacle_counts[b] rows (counts >= n_obstacle, as
 * np.random.choice(replace=False) requires; smaller counts wrap around instead of raising). */
int mpn_build_cloud_from_points(mpn_ctx* ctx, void* stream, int B, const float* q0, const float* target,
                                const float* obstacle_points, const int32_t* obstacle_counts, int max_points,
                                uint32_t problem0, float* cloud);
/* Depth-camera obstacle clouds: stand-in for run_inference.convert_primitive_problems_to_depth (run_inference.py:194-257),
 * which renders the primitives (robot removed) with Bullet from the fixed evaluation cameras of :215-243 and un-projects the
 * depth image (robofin.bullet.get_pointcloud_from_camera, un-vendored).  Here each pixel's ray is intersected analytically
 * with every valid cuboid / cylinder; the nearest hit with depth in [near_depth, far_depth] gives one world point.
 *   camera: camera->world pose(s), 3x4 row-major, [12] shared or [B][12] (per_problem_camera != 0); OpenGL camera frame
 *   (x right, y up, looking along -z: the convention under which the poses of run_inference.py:215-243 face their scenes);
 *   pixel (row, col), row 0 at the top, looks along ((2 (col + .5) / W - 1) tan_half_fov_x, -(2 (row + .5) / H - 1)
 *   tan_half_fov_y, -1).  points [B][W*H][3]: the hits compacted to the front in pixel order; counts i32 [B].
 * The result feeds mpn_build_cloud_from_points. */
int mpn_render_depth_cloud(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, const float* camera,
                           int per_problem_camera, int width, int height, float tan_half_fov_x, float tan_half_fov_y,
                           float near_depth, float far_depth, float* points, int32_t* counts);
/* validation collision sweep (model.py:293-314): traj [B][T][7] unnormalised -> flags u8 [B] (OR-ed into existing
 * content when accumulate != 0), first_step i32 [B] (optional; step index offset by t0; -1 when none) */
int mpn_sweep_flags(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, const float* traj, int T, int t0,
                    int accumulate, uint8_t* flags, int32_t* first_step);

/* ---- mpinets.model replacements ----------------------------------------------------------------------------- */
/* MPiNetsPointNet.forward (model.py:409-426): cloud [B][N][4] -> [B][2048] */
int mpn_encoder_forward(mpn_ctx* ctx, void* stream, int precision, const float* cloud, int B, int N, float* out);
/* MotionPolicyNetwork.forward (model.py:75-91): cloud [B][N][4], q_norm [B][7] -> delta q [B][7] */
int mpn_policy_forward(mpn_ctx* ctx, void* stream, int precision, const float* cloud, const float* q_norm, int B,
                       int N, float* dq);

#define MPN_METRICS_COLS 8
/* columns of the metrics table: */
enum { MPN_M_COLLISION = 0, MPN_M_FIRST_COLLISION_STEP = 1, MPN_M_STEPS = 2, MPN_M_POS_ERR = 3, MPN_M_ORI_ERR_DEG = 4,
       MPN_M_REACHED = 5, MPN_M_MIN_SDF_MARGIN = 6, MPN_M_RESERVED = 7 };

/* TrainingMotionPolicyNetwork.rollout + validation sweep (model.py:128-183,272-314) and, with early_exit != 0,
 * run_inference.rollout_until_success (run_inference.py:137-191) in lock-step with a per-problem done mask.
 * cloud [B][N][4] is updated in place (robot rows), like the reference (model.py:181).
 * q0 [B][7] unnormalised start; target [B][12]; traj [B][T+1][7] unnormalised (incl. start); metrics [B][8] fp32.
 * check_every_step != 0 evaluates the collision flag after every step (config 3) instead of once at the end.
 * early_exit: 0 = exactly T steps for everybody; 1 = per-problem done mask (stopped problems keep their configuration) and the host
 * polls every 8 steps whether every problem has stopped, ending the loop early like rollout_until_success's break (the only host
 * synchronisation of this call; skipped while the stream is being captured) and, once at most half of the current problems are still
 * running, carries only those on in a compact set (results scattered back; a stopped problem's rows of `cloud` keep the robot points of
 * the last step it was carried through); 2 = done mask only, never synchronises. */
int mpn_rollout(mpn_ctx* ctx, void* stream, int precision, const mpn_scene* scene, int B, int N, float* cloud,
                const float* q0, const float* target, int T, int early_exit, int check_every_step, float* traj,
                float* metrics);

/* Evaluator.evaluate_trajectory, the device-computable subset (metrics.py:311-322 joint limits, :340-384 final position /
 * orientation / region, :411-434 end-effector path lengths, :487-523 the success conjunction), for B trajectories at once.
 *   traj [B][n_poses_max][7] unnormalised (mpn_rollout's layout with n_poses_max = T+1); num_poses i32 [B] (optional:
 *   valid poses per trajectory, e.g. steps+1 after an early exit; NULL = all); target [B][12] right_gripper pose.
 *   target_volume / negative_volumes: per-problem primitive lists in mpn_scene layout with their own row counts
 *   ([B][tv_cuboids][..] etc.; zero-volume rows = padding; either may be NULL with counts 0 -> region test passes).
 *   They are geometrout primitives in the reference, so their frames use textbook rotations regardless of quirk_frames.
 * Differences from the reference, by construction: "collision" is the validation sphere sweep (model.py:293-314) instead of
 * Bullet; "self_collision" is a sphere-sphere test between collision spheres whose link groups (link1..6, {link7, hand,
 * fingers}) are >= 2 apart, standing in for robofin's FrankaSelfCollisionChecker; SPARC smoothness is not computed. */
#define MPN_EVAL_COLS 16
enum { MPN_E_COLLISION = 0, MPN_E_JOINT_LIMIT_VIOLATION = 1, MPN_E_SELF_COLLISION = 2, MPN_E_PHYSICAL_VIOLATIONS = 3,
       MPN_E_POSITION_ERROR_CM = 4, MPN_E_ORIENTATION_ERROR_DEG = 5, MPN_E_EFF_POSITION_PATH_LENGTH = 6,
       MPN_E_EFF_ORIENTATION_PATH_LENGTH_DEG = 7, MPN_E_CORRECT_FINAL_REGION = 8, MPN_E_SUCCESS = 9, MPN_E_NUM_STEPS = 10,
       MPN_E_FIRST_COLLISION_STEP = 11, MPN_E_CONFIG_PATH_LENGTH = 12, MPN_E_MAX_COLLISION_DEPTH = 13 };
int mpn_evaluate(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, const float* traj, int n_poses_max,
                 const int32_t* num_poses, const float* target, const mpn_scene* target_volume, int tv_cuboids,
                 int tv_cylinders, const mpn_scene* negative_volumes, int nv_cuboids, int nv_cylinders, float* eval);

/* third_party/sparc.py:48-140 (spectral arc length; Evaluator.calculate_smoothness, metrics.py:387-409) for B speed profiles:
 * movement [B][n_max] (num_samples i32 [B] optional: valid prefix per row), fs sampling frequency; the reference's
 * defaults are padlevel 4, fc 10.0, amp_th 0.05.  sal [B]; 0 for an all-zero profile (sparc.py:95-97). */
int mpn_sparc(mpn_ctx* ctx, void* stream, int B, int n_max, const float* movement, const int32_t* num_samples, float fs,
              int padlevel, float fc, float amp_th, float* sal);

/* ---- training losses (mpinets/loss.py), forward value + analytic gradient; all pointers are device pointers ----------
 * collision_loss (loss.py:47-94): loss[0] = mean over B*N of max(0, margin - sdf(points[b][n])) with the scene sdf of
 * geometry.py:238-288,456-507 (F.hinge_embedding_loss with target -1; the reference uses margin 0.03).
 * grad_points (optional, [B][N][3]) = d loss / d points with torch autograd's conventions. */
int mpn_collision_loss(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, int N, const float* points, float margin,
                       float* loss, float* grad_points);
/* point_match_loss (loss.py:31-44): loss[0] = mse_mean(a, b) + l1_mean(a, b) over n floats; grad_a optional [n] */
int mpn_point_match_loss(mpn_ctx* ctx, void* stream, int64_t n, const float* a, const float* b, float* loss, float* grad_a);
/* CollisionAndBCLossContainer.__call__ (loss.py:111-166): input_normalized / target_normalized [B][7] in [-1, 1] ->
 * unnormalise -> FK -> the fixed n_points robot cloud (FrankaSampler(num_fixed_points = n_points, with_base_link = False):
 * here the first n_points entries of a seeded permutation of the link table's non-base rows) -> losses[0] = collision
 * loss, losses[1] = point-match loss.  grad_input (optional, [B][7]) = d(w_collision * losses[0] + w_bc * losses[1]) /
 * d input_normalized (model.py:232-236 uses the weights of jobconfig.yaml:24-25). */
int mpn_bc_collision_losses(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, const float* input_normalized,
                            const float* target_normalized, int n_points, float margin, float w_collision, float w_bc,
                            float* losses, float* grad_input);

/* ---- training step (TrainingMotionPolicyNetwork.training_step, model.py:185-240) ------------------------------------------
 * Parameters live in ONE flat fp32 device vector in state-dict order (SA_modules.{0,1,2}.mlps.0.{0,2,4}.{weight,bias},
 * fc_layer.{0,1,3,4,6}, feature_encoder.{0..8}, decoder.{0..6}), every tensor starting at a multiple of 4 floats;
 * gradients use the same layout, so the host all-reduces one vector (DDP, run_training.py:71-77) and the optimiser is one
 * pass.  mpn_param_info enumerates the tensors (returns MPN_ERR_INVALID past the last index). */
int64_t mpn_param_count(mpn_ctx* ctx);
int mpn_param_info(mpn_ctx* ctx, int index, char* name, int name_cap, int64_t* offset, int64_t* numel);
int mpn_get_params(mpn_ctx* ctx, void* stream, float* dst /* device [param_count] */);
int mpn_set_params(mpn_ctx* ctx, void* stream, const float* src /* device [param_count] */);
/* rebuilds the packed bf16 tensor-core copies from the fp32 parameters (after optimisation, before MPN_PREC_BF16 inference);
 * synchronises the device */
int mpn_weights_sync(mpn_ctx* ctx);
/* forward (fp32 kernels, state saved for the backward pass) -> y_hat = clamp(q_norm + net(cloud, q_norm), -1, 1) (model.py:202) ->
 * losses[0] = collision loss, losses[1] = point-match loss of CollisionAndBCLossContainer (loss.py:111-166) against
 * `supervision` [B][7] -> grads [param_count] = d(w_collision * losses[0] + w_bc * losses[1]) / d parameters (overwritten;
 * NULL: forward + losses only).  cloud [B][N][4], q_norm [B][7] in [-1, 1]; y_hat [B][7] optional output.
 * The reference's values: n_loss_points 1024, margin 0.03 (loss.py:92,109), w_collision 5, w_bc 1 (jobconfig.yaml:24-25).
 * precision: MPN_PREC_FP32 = forward and every GEMM of the backward in fp32 (the parity mode); MPN_PREC_BF16 = the counterpart of
 * the reference's precision=16 autocast (run_training.py:109): the point-cloud encoder's forward runs through the fused
 * tensor-core kernels (which also record the max-pool routing), the compacted-row GEMMs of the SA1 / SA2 backward and the
 * data-gradient GEMMs of the group-all level run on tcgen05 with bf16 operands and fp32 accumulation; master weights,
 * optimizer state and parameter gradients are fp32 in both.  mpn_adam_step refreshes the packed bf16 copies on its stream. */
int mpn_train_step_grads(mpn_ctx* ctx, void* stream, const mpn_scene* scene, int B, int N, const float* cloud,
                         const float* q_norm, const float* supervision, int n_loss_points, float margin, float w_collision,
                         float w_bc, float* losses, float* y_hat, float* grads, int precision);
/* Tensor-core building blocks of the training backward (train_tc.cu), exposed for their own parity tests.
 * mpn_train_tc_gemm: C[M][N] bf16 = epi(A[M][128] bf16 * W[N][128]^T bf16 + bias), M % 128 == 0, N in {64, 128, 256};
 *   epi 0 ReLU, 1 none, 2 no bias, multiplied by relu'(mask[M][N] bf16) (C may alias mask).
 * mpn_train_tc_wgrad: partial[n_ctas][128][128] fp32 = per-CTA sums over row ranges of dY[r][:]^T X[r][:] (dY, X [R][128]
 *   bf16 row-major, handed to tcgen05 as MN-major operands); the sum over the n_ctas tiles is dY^T X.  variant 0. */
int mpn_train_tc_gemm(mpn_ctx* ctx, void* stream, int epi, const void* A, const void* W, const float* bias, const void* mask,
                      int64_t M, int N, void* C);
int mpn_train_tc_wgrad(mpn_ctx* ctx, void* stream, const void* dY, const void* X, int64_t R, float* partial,
                       int64_t partial_floats, int* n_ctas, int variant);
/* the max-pool routing of the last training step: for module 0 / 1 / 2 the neighbour row (0..127, in ball-query order;
 * group-all: the SA2 centroid) that won each output channel, u8 [B][512][64] / [B][128][256] / [B][1024] (device).  The
 * backward pass sends each channel's gradient to exactly this row (max_pool2d backward); parity tests replay it. */
int mpn_train_pooled_rows(mpn_ctx* ctx, void* stream, int module, int B, uint8_t* dst);
/* torch.nn.utils.clip_grad_norm_(max_norm = clip_norm; <= 0 disables; run_training.py:112 uses 1.0) followed by
 * torch.optim.Adam (model.py:68-73: lr 1e-4, betas (0.9, 0.999), eps 1e-8) on the flat vector; `step` counts from 1;
 * grad_norm (optional, device float) receives the total gradient norm before clipping. */
int mpn_adam_step(mpn_ctx* ctx, void* stream, const float* grads, float lr, float beta1, float beta2, float eps,
                  float clip_norm, int step, float* grad_norm);

/* number of kernels this library has launched on this context since creation (bench.py's gpu_launches) */
int64_t mpn_launch_count(mpn_ctx* ctx);

/* per-stage device timing with CUDA events recorded on the launching stream (bench.py's roofline leg).
 * mpn_profile(ctx, 1) starts recording, mpn_profile_read synchronises the device and returns, for each stage,
 * the accumulated milliseconds and the number of timed launches since the last read. */
#define MPN_NUM_STAGES 12
enum { MPN_ST_FPS1 = 0, MPN_ST_SA1 = 1, MPN_ST_FPS2 = 2, MPN_ST_SA2 = 3, MPN_ST_SA3 = 4, MPN_ST_FC = 5, MPN_ST_HEADS = 6,
       MPN_ST_UPDATE = 7, MPN_ST_SAMPLE_ROBOT = 8, MPN_ST_SWEEP = 9, MPN_ST_BUILD_CLOUD = 10, MPN_ST_OTHER = 11 };
int mpn_profile(mpn_ctx* ctx, int enable);
int mpn_profile_read(mpn_ctx* ctx, float* ms /*[MPN_NUM_STAGES]*/, int64_t* launches /*[MPN_NUM_STAGES]*/);

/* tensor-core self-test: D[128][N] fp32 = A[128][K] bf16 * B[N][K]^T bf16 on one CTA with the smem-descriptor
 * convention `mode` (0: interleaved LBO=K-dir; 1: interleaved, LBO/SBO swapped; 2: 128B swizzle; 3: interleaved MN-first).
 * status (device int) is set to 1 when the MMA completion barrier timed out. */
/* synchronises and returns the tensor-core path's sticky error flag (1 = an MMA completion barrier timed out) */
int mpn_tc_error(mpn_ctx* ctx, int* out);
/* executed-work accounting of the tensor-core set-abstraction kernels (replaces nothing in the reference; bench.py's roofline leg):
 * the fused SA kernels pack the DISTINCT neighbour rows of several groups into shared 128-row MMA tiles (pointnet2's ball query pads a
 * group with copies of its first hit -- ball_query_gpu.cu semantics, SURVEY App. A.1 -- and the max-pool ignores duplicates), so the
 * tensor work they issue is (tiles) x (MMA flops per tile), not the reference's 128 rows per group.  out[0] / out[1] = number of tiles
 * issued by the SA1 / SA2 kernels of this context since the last reset; synchronises the device. */
int mpn_sa_tile_counts(mpn_ctx* ctx, uint64_t* out /*[2]*/, int reset);
int mpn_tc_selftest(mpn_ctx* ctx, void* stream, const void* a_bf16, const void* b_bf16, float* d, int N, int K, int mode,
                    int* status);

/* GEMM self-test of the TMA / tcgen05 row GEMM (gemm_tc.cu) behind SA3 and the FC head: out[M][N] fp32 = a[M][K] * w[N][K]^T +
 * bias, operands taken as bf16 (split = 0) or as split-bf16 (hi, lo) pairs with the three-pass product of MPN_PREC_BF16X3
 * (split = 1).  Test entry: allocates temporaries and synchronises the stream. */
int mpn_tc_gemm_selftest(mpn_ctx* ctx, void* stream, const float* a, const float* w, const float* bias, int M, int N, int K,
                         float* out, int split);

#ifdef __cplusplus
}
#endif
#endif /* MPINETS_B200_H */
