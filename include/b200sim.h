/*
 * b200sim.h -- C-ABI of libb200sim.so: the B200-native batched rigid-body backend that sits where the
 * reference (haosulab/ManiSkill) calls into `sapien.physx.PhysxGpuSystem`.
 *
 * Plain C linkage, plain pointers and sizes, no torch / C++ types.  Every function returns 0 on success and a
 * negative B2S_ERR_* code on failure; b2s_last_error() returns a human readable message for the calling thread.
 *
 * Each entry point names the reference call site (file:line under /root/reference/mani_skill) whose backend
 * call it replaces:
 *
 *   b2s_world_create          PhysxGpuSystem(device) + N x sapien.Scene + builders' .build()   envs/sapien_env.py:1186-1210
 *   b2s_world_buffers         px.gpu_init() + px.cuda_* buffer objects (.torch())               envs/scene.py:902-948
 *   b2s_step                  px.step()                                                         envs/scene.py:379-380 (loop at envs/sapien_env.py:1123-1128)
 *   b2s_apply                 px.gpu_apply_{rigid_dynamic_data,articulation_qpos,qvel,qf,root_pose,root_velocity,
 *                             target_position,target_velocity}()                               envs/scene.py:950-966, envs/sapien_env.py:1118-1121
 *   b2s_fetch                 px.gpu_fetch_{rigid_dynamic_data,articulation_link_pose,link_velocity,qpos,qvel,qacc,
 *                             target_qpos,target_qvel}()                                       envs/scene.py:968-986
 *   b2s_update_kinematics     px.gpu_update_articulation_kinematics()                           envs/scene.py:947, envs/sapien_env.py:958
 *   b2s_contact_query_create  px.gpu_create_contact_pair_impulse_query(body_pairs)              envs/scene.py:761-775
 *   b2s_contact_query_run     px.gpu_query_contact_pair_impulses(query)                         envs/scene.py:776-781
 *                             (row b = B2S_ANY_BODY: px.gpu_create_contact_body_impulse_query /
 *                             px.gpu_query_contact_body_impulses, the net impulse on body a)         utils/structs/base.py:116-136, articulation.py:441-462
 *   b2s_camera_group_create   RenderSystemGroup.create_camera_group(cameras, texture_names)     envs/scene.py:1087-1106
 *   b2s_render                camera_group.take_picture() (+ set_cuda_poses / update_render)    utils/structs/render_camera.py:269-273, envs/scene.py:404-427
 *   b2s_pick_task_create/step BaseEnv.step() for the PickCube-v1 family (controller + 5 substeps + evaluate + obs +
 *                             reward) as one launch sequence                                   envs/sapien_env.py:1042-1132
 *   b2s_pick_task_step_autoreset  the same + ManiSkillVectorEnv's auto-reset of finished sub-scenes on the device
 *                                                                                               vector/wrappers/gymnasium.py:160-176
 *
 * Memory: all device buffers are owned by the world (cudaMalloc at create); b2s_world_buffers() hands out
 * borrowed device pointers that alias the live state, exactly like `px.cuda_rigid_body_data.torch()` does.
 * All launches go to the cudaStream_t passed in (as void*); 0 = legacy default stream.
 */
#ifndef B200SIM_H_
#define B200SIM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_OK 0
#define B2S_ERR_INVALID -1
#define B2S_ERR_CUDA -2
#define B2S_ERR_CAPACITY -3
#define B2S_ERR_NO_DEVICE -4

/* shape types */
#define B2S_SHAPE_PLANE 0   /* half-space, normal = +x of the shape pose (sapien convention) */
#define B2S_SHAPE_BOX 1     /* size = half extents */
#define B2S_SHAPE_SPHERE 2  /* size[0] = radius */
#define B2S_SHAPE_CAPSULE 3 /* size[0] = radius, size[1] = half length along local x */
#define B2S_SHAPE_CONVEX 4  /* hull id, vertices in hull_verts */
/* shape owner kinds */
#define B2S_OWNER_STATIC 0 /* pose is in the sub-scene frame */
#define B2S_OWNER_LINK 1   /* owner = dof/abody index, or -(art+1) for links rigidly attached to a fixed root */
#define B2S_OWNER_BODY 2   /* owner = free body index */
/* joint types of moving (1-dof) joints */
#define B2S_JOINT_REVOLUTE 0
#define B2S_JOINT_PRISMATIC 1
/* free body types */
#define B2S_BODY_DYNAMIC 0
#define B2S_BODY_KINEMATIC 1

/*
 * Compiled scene model: ONE sub-scene prototype instantiated n_envs times (env-major struct-of-arrays on the
 * device).  Produced by maniskill_b200/model.py from ManiSkill-style builders; all pointers are HOST memory and
 * are copied during b2s_world_create.  Layout mirrors what ManiSkill hands to sapien at build time
 * (utils/building/actor_builder.py:57-164, utils/building/articulation_builder.py:65-112).
 *
 * Articulations are fixed-base trees of 1-dof joints (revolute / prismatic); links joined by fixed joints are
 * merged into one "abody" per moving joint for dynamics (their poses are still reported per link).
 */
typedef struct B2SModel {
  int32_t n_envs;
  int32_t n_art;        /* articulations per env */
  int32_t n_dof;        /* moving joints per env (all articulations), <= 32 */
  int32_t n_link;       /* links per env (all articulations) */
  int32_t n_fb;         /* free rigid bodies (dynamic + kinematic actors) per env */
  int32_t n_shape;      /* collision shapes per env */
  int32_t n_pair;       /* candidate shape pairs per env (broadphase list, already group/SRDF filtered) */
  int32_t n_hull;
  int32_t n_hull_verts; /* total vertices over all hulls */
  int32_t n_eq;         /* joint coupling rows (fixed tendons / mimic joints) */
  int32_t n_ov_shape;   /* shapes with per-env size+pose overrides (heterogeneous envs) */
  int32_t n_ov_fb;      /* free bodies with per-env mass properties */
  int32_t max_contacts; /* contact point capacity per env */
  int32_t max_manifolds;/* contact patch (shape pair) capacity per env */
  int32_t n_pos_iters;  /* SceneConfig.solver_position_iterations (utils/structs/types.py:46) */
  int32_t n_vel_iters;  /* SceneConfig.solver_velocity_iterations (utils/structs/types.py:47) */
  int32_t max_dof_per_art;
  float dt;             /* 1 / sim_freq (envs/sapien_env.py:1227) */
  float gravity[3];
  float contact_offset; /* SceneConfig.contact_offset (types.py:44) */
  float rest_offset;
  float max_depen_vel;  /* cap on penetration recovery speed */
  float contact_hertz;  /* soft-contact natural frequency for penetration recovery (Hz) */
  float contact_zeta;   /* soft-contact damping ratio */
  float margin_min;     /* speculative contact margin for a pair at rest (grows with approach speed up to 2*contact_offset) */
  /* ---- moving joints / abodies, [n_dof] ---- */
  const int32_t* dof_parent;   /* parent abody or -1 (child of the fixed root) */
  const int32_t* dof_art;
  const int32_t* dof_type;
  const float* dof_T0;         /* [n_dof*7] joint frame in parent abody frame: p(3) q(wxyz) */
  const float* dof_axis;       /* [n_dof*3] unit axis in the joint frame */
  const float* dof_mass;
  const float* dof_com;        /* [n_dof*3] in abody frame */
  const float* dof_inertia;    /* [n_dof*6] xx yy zz xy xz yz about com, abody axes */
  const float* dof_gravity;    /* [n_dof] 1 = gravity acts, 0 = disabled (agents/base_agent.py:278-282) */
  const float* dof_limit;      /* [n_dof*2] lower, upper (+-1e30 = none) */
  const float* dof_drive;      /* [n_dof*4] stiffness, damping, force_limit, max joint velocity (0 = unlimited; PhysX default 100) */
  const float* dof_passive;    /* [n_dof*4] joint damping, joint friction, armature, reserved */
  const uint32_t* dof_anc_mask;/* [n_dof] bit j set iff abody j is an ancestor-or-self */
  /* ---- links, [n_link] ---- */
  const int32_t* link_dof;     /* abody the link is rigidly part of, or -(art+1) if fixed to the root */
  const float* link_offset;    /* [n_link*7] link frame in abody (or root) frame */
  /* ---- articulation roots, [n_art] ---- */
  const float* art_root_pose;  /* [n_art*7] initial root pose in the sub-scene frame */
  const int32_t* art_dof_start;/* [n_art+1] */
  const int32_t* art_link_start;/* [n_art+1] */
  /* ---- coupling rows ---- */
  const int32_t* eq_dof;       /* [n_eq*2] (a, b): q_b - mult*q_a - offset = 0 */
  const float* eq_param;       /* [n_eq*4] mult, offset, stiffness, reserved */
  /* ---- free bodies, [n_fb] ---- */
  const int32_t* fb_type;
  const float* fb_mass;
  const float* fb_com;         /* [n_fb*3] */
  const float* fb_inertia;     /* [n_fb*6] */
  const float* fb_damping;     /* [n_fb*2] linear, angular */
  const float* fb_gravity;     /* [n_fb] */
  const float* fb_init_pose;   /* [n_fb*7] */
  const int32_t* fb_ov;        /* [n_fb] per-env mass-property slot or -1 */
  /* ---- shapes, [n_shape] ---- */
  const int32_t* shape_type;
  const int32_t* shape_owner_kind;
  const int32_t* shape_owner;
  const int32_t* shape_row;    /* exposed body row (link / actor) the shape belongs to, -1 for static */
  const float* shape_pose;     /* [n_shape*7] in owner frame */
  const float* shape_size;     /* [n_shape*3] */
  const int32_t* shape_hull;
  const float* shape_mu;       /* friction coefficient (static == dynamic) */
  const float* shape_bound;    /* [n_shape*4] bounding sphere centre (owner frame) + radius */
  const int32_t* shape_ov;     /* per-env override slot or -1 */
  const float* shape_patch;    /* min torsional patch radius (urdf_config patch_radius, agents/robots/panda/panda.py:24-31) */
  /* ---- hulls ---- */
  const int32_t* hull_offset;  /* [n_hull+1] */
  const float* hull_verts;     /* [n_hull_verts*3] */
  const float* hull_aabb;      /* [n_hull*6] local box (centre, half extents) of each hull, broadphase only */
  /* ---- broadphase candidates ---- */
  const int32_t* pair_a;
  const int32_t* pair_b;
  /* ---- per-env overrides (host arrays, env-major) ---- */
  const float* ov_shape_size;  /* [n_envs*n_ov_shape*3] */
  const float* ov_shape_pose;  /* [n_envs*n_ov_shape*7] */
  const float* ov_shape_bound; /* [n_envs*n_ov_shape*4] */
  const float* ov_fb_mass;     /* [n_envs*n_ov_fb*10] mass, com(3), inertia(6) */
} B2SModel;

/* Device views returned by b2s_world_buffers (all float32 unless noted; env-major AoS like the reference's
 * px.cuda_* buffers, utils/structs/base.py:262-270, articulation.py:723-815). */
typedef struct B2SBufferTable {
  float* rigid_body_data;  /* [n_envs*(n_link+n_fb), 13]  pos3 quat(wxyz)4 linvel3 angvel3; row = env*(rows)+r */
  float* qpos;             /* [n_envs*n_art, max_dof_per_art] */
  float* qvel;
  float* qacc;
  float* qf;
  float* target_qpos;
  float* target_qvel;
  int32_t n_rows;          /* rigid rows per env = n_link + n_fb */
  int32_t max_dof;
  int32_t* contact_count;  /* [n_envs] contacts generated in the last substep */
  int32_t* overflow_flag;  /* [1] sticky OR of B2S_OVF_* reasons: a fixed capacity dropped a contact or constraint row in some sub-scene */
} B2SBufferTable;

/* bits of *overflow_flag (PhysX reports exceeded GPUMemoryConfig capacities, mani_skill/utils/structs/types.py:16-32) */
#define B2S_OVF_MANIFOLDS 1  /* more contact patches than max_manifolds (<= 24) in a sub-scene */
#define B2S_OVF_CONTACTS 2   /* more contact points than max_contacts (<= 64) */
#define B2S_OVF_ROWS 4       /* more constraint rows than the compiled row capacity */
#define B2S_OVF_ART_ROWS 8   /* (unused since rows touching an articulation share the row capacity) */
#define B2S_OVF_LIMITS 16    /* more simultaneously active joint limits than 24 */
#define B2S_OVF_EQ 32        /* more tendon couplings than 2 */

/* apply / fetch selection bits */
#define B2S_BUF_RIGID (1u << 0)      /* free-body rows of rigid_body_data (pose + velocity) */
#define B2S_BUF_ROOT_POSE (1u << 1)  /* articulation root rows (root link row) */
#define B2S_BUF_QPOS (1u << 2)
#define B2S_BUF_QVEL (1u << 3)
#define B2S_BUF_QF (1u << 4)
#define B2S_BUF_TARGET_QPOS (1u << 5)
#define B2S_BUF_TARGET_QVEL (1u << 6)
#define B2S_BUF_QACC (1u << 7)
#define B2S_BUF_LINK (1u << 8)       /* link pose + velocity rows (fetch only) */
#define B2S_BUF_ALL 0xFFFFFFFFu

const char* b2s_last_error(void);
int32_t b2s_version(void);

int32_t b2s_world_create(const B2SModel* model, int32_t device, uint64_t* world);
int32_t b2s_world_destroy(uint64_t world);
int32_t b2s_world_buffers(uint64_t world, B2SBufferTable* out);

/* Advance all sub-scenes by `substeps` physics steps of model.dt (fused: PD drive, ABA, collide, TGS, integrate).
 * fetch_mask != 0 additionally refreshes the exposed buffers after the last substep (fused gpu_fetch_*). */
int32_t b2s_step(uint64_t world, int32_t substeps, uint32_t fetch_mask, void* stream);
int32_t b2s_apply(uint64_t world, uint32_t mask, void* stream);
int32_t b2s_fetch(uint64_t world, uint32_t mask, void* stream);
int32_t b2s_update_kinematics(uint64_t world, void* stream);

/* Sum of last-substep solver contact impulses (sub-scene frame) between exposed body rows a and b, acting on a.
 * rows: [n_query*2] per-env row ids (-1 = the static geometry; b = B2S_ANY_BODY sums over every body touching a: the net
 * contact impulse on a).  out: device [n_envs, n_query, 3]. */
#define B2S_ANY_BODY (-2)
int32_t b2s_contact_query_create(uint64_t world, const int32_t* rows, int32_t n_query, uint64_t* query);
int32_t b2s_contact_query_run(uint64_t world, uint64_t query, float* out_dev, void* stream);

/* Camera group: n_cam cameras per env, each w x h.  Render primitives come from the model's visual table. */
typedef struct B2SCameraDesc {
  int32_t width, height;
  float fx, fy, cx, cy, near_, far_;
  int32_t mount_row;   /* exposed body row the camera is mounted on, or -1 for sub-scene frame */
  float local_pose[7]; /* camera pose (sapien convention: x forward, y left, z up) in mount frame */
} B2SCameraDesc;

typedef struct B2SVisualTable {
  int32_t n_visual;        /* <= 64 render shapes per sub-scene */
  const int32_t* type;     /* B2S_SHAPE_* (convex hulls and boxes are rasterised, spheres / planes / near-plane crossing boxes ray-cast) */
  const int32_t* row;      /* exposed body row the shape follows, or -1 static */
  const float* pose;       /* [n*7] local pose in the body frame */
  const float* size;       /* [n*3] */
  const float* color;      /* [n*4] base colour rgba in [0,1] */
  const int32_t* seg_id;   /* per_scene_id of the owning entity (segmentation value) */
  const int32_t* ov_slot;  /* [n] per-env override slot or -1 */
  int32_t n_ov;
  const float* ov_size;    /* [n_envs*n_ov*3] env-major */
  const float* ov_pose;    /* [n_envs*n_ov*7] env-major */
  /* indexed triangle geometry of the visuals that are rasterised: convex hulls (vertices in the visual's frame) and boxes
   * (the 8 corners of the unit cube (+-1), scaled by the -- possibly per-env -- half extents when they are projected) */
  int32_t n_vert;
  const float* vert_local;  /* [n_vert*3] */
  const int32_t* vert_vis;  /* [n_vert] owning visual */
  int32_t n_tri;
  const int32_t* tri_idx;   /* [n_tri*3] vertex indices, wound so that the normal points out of the solid */
  const int32_t* tri_vis;   /* [n_tri] owning visual */
} B2SVisualTable;

/* what a camera group writes per pixel: the shader pack's raw render targets (render/shaders.py:68-84, what
 * `get_picture_cuda(name)` hands out) and / or the textures the observation modes deliver (render/shaders.py:74-83 texture_transforms:
 * rgb = Color[..., :3], depth = -PositionSegmentation[..., 2], segmentation = PositionSegmentation[..., 3]) as contiguous tensors */
#define B2S_OUT_COLOR 1u
#define B2S_OUT_POSSEG 2u
#define B2S_OUT_RGB 4u
#define B2S_OUT_DEPTH 8u
#define B2S_OUT_SEG 16u
#define B2S_OUT_RAW (B2S_OUT_COLOR | B2S_OUT_POSSEG)

typedef struct B2SRenderTargets {
  uint8_t* color;        /* [n_envs, n_cam, h, w, 4] rgba8  (render/shaders.py:68-84 "Color"); NULL unless B2S_OUT_COLOR */
  int16_t* position_seg; /* [n_envs, n_cam, h, w, 4] int16: x,y,z (mm, OpenGL camera frame), segmentation id; NULL unless B2S_OUT_POSSEG */
  uint8_t* rgb;          /* [n_envs, n_cam, h, w, 3] uint8; NULL unless B2S_OUT_RGB */
  int16_t* depth;        /* [n_envs, n_cam, h, w] int16 mm; NULL unless B2S_OUT_DEPTH */
  int16_t* segmentation; /* [n_envs, n_cam, h, w] int16; NULL unless B2S_OUT_SEG */
} B2SRenderTargets;

/* raw render targets (B2S_OUT_RAW) */
int32_t b2s_camera_group_create(uint64_t world, const B2SCameraDesc* cams, int32_t n_cam, const B2SVisualTable* vis,
                                uint64_t* group, B2SRenderTargets* out);
/* the same with a choice of outputs (a mask of B2S_OUT_*) */
int32_t b2s_camera_group_create_outputs(uint64_t world, const B2SCameraDesc* cams, int32_t n_cam, const B2SVisualTable* vis,
                                        uint32_t outputs, uint64_t* group, B2SRenderTargets* out);
int32_t b2s_render(uint64_t world, uint64_t group, void* stream);

/*
 * Fused control step for the pick-and-place task family (PickCube-v1): what BaseEnv.step() does between receiving the
 * action and returning (obs, reward, terminated, truncated, info) -- envs/sapien_env.py:1042-1132 -- as three launches:
 *   1. joint-space controller: clip/scale the normalised action, delta or absolute targets, mimic joints
 *      (agents/controllers/pd_joint_pos.py:76-93,207-228, utils/gym_utils.py:104-108) -> target_qpos
 *   2. `substeps` physics steps + fetch (b2s_step)
 *   3. epilogue: is_grasped from the finger/object contact impulses (agents/robots/panda/panda.py:237-265), is_obj_placed,
 *      is_robot_static, success (envs/tasks/tabletop/pick_cube.py:147-159), dense reward (:161-191), flattened state
 *      observation (:132-145 + agents/base_agent.py:339-347), elapsed_steps += 1, truncation (utils/registration.py:160-168)
 * The python path (maniskill_b200/envs) computes the same values with torch ops and is kept as the checked fallback.
 */
typedef struct B2SJointController {
  int32_t n_action;             /* action width */
  const int32_t* dof_action;    /* [n_dof] action column driving this dof, or -1 (target untouched) */
  const int32_t* dof_use_delta; /* [n_dof] 1: target = qpos + scaled action, 0: target = scaled action */
  const int32_t* dof_normalize; /* [n_dof] 1: clip to [-1,1] and scale into [low, high] */
  const float* dof_low;         /* [n_dof] */
  const float* dof_high;        /* [n_dof] */
} B2SJointController;

typedef struct B2SPickTask {
  int32_t tcp_row, obj_row, goal_row, lfinger_row, rfinger_row; /* exposed body rows */
  float goal_thresh;     /* pick_cube.py:43 */
  float min_force;       /* panda.py:237 (0.5 N) */
  float max_angle_deg;   /* panda.py:237 (85) */
  float static_thresh;   /* pick_cube.py:154 (0.2) */
  int32_t n_static_dof;  /* leading dofs checked by is_static / static reward (qvel[..., :-2]) */
  int32_t max_episode_steps;
  int32_t normalized_reward; /* 1: reward / 5 */
} B2SPickTask;

typedef struct B2SPickOutputs {
  float* obs;          /* [n_envs, 2*n_dof + 24] : qpos, qvel, is_grasped, tcp_pose7, goal_pos3, obj_pose7, tcp_to_obj3, obj_to_goal3 */
  float* reward;       /* [n_envs] */
  uint8_t* flags;      /* [n_envs, 6] : success, is_obj_placed, is_robot_static, is_grasped, terminated, truncated */
  int32_t* elapsed;    /* [n_envs] in/out */
} B2SPickOutputs;

int32_t b2s_pick_task_create(uint64_t world, const B2SJointController* ctrl, const B2SPickTask* task, uint64_t* handle);
/* actions: device [n_envs, n_action] float32, or NULL to step without a new action */
int32_t b2s_pick_task_step(uint64_t world, uint64_t handle, const float* actions_dev, int32_t substeps, const B2SPickOutputs* out,
                           void* stream);

/*
 * Auto-reset on the device -- `ManiSkillVectorEnv.step` (vector/wrappers/gymnasium.py:127-184: final_info / final_observation
 * bookkeeping, partial `reset(options={"env_idx": ...})`) + `BaseEnv.reset` for the sub-scenes that finished
 * (envs/sapien_env.py:857-978 -> PickCube `_initialize_episode`, tasks/tabletop/pick_cube.py:106-130, and
 * `TableSceneBuilder.initialize`, utils/scene_builder/table/scene_builder.py:68-103) without the `dones.any()` host sync:
 * after the control step the kernels (1) mark done = terminated | truncated (elapsed >= max_episode_steps), copy the observation of
 * done sub-scenes to `final_obs`, (2) re-initialise their state from the caller's random numbers (cube xy / yaw, goal xyz uniform,
 * robot rest pose + Gaussian noise, zero velocities, drive targets = reset pose, elapsed = 0), (3) refresh the exposed buffers and
 * (4) rewrite their observation row.  Which random stream feeds an un-seeded reset is unspecified in the reference (the global torch
 * RNG); here it is the `rand` tensor the caller fills each step.
 */
typedef struct B2SPickReset {
  float cube_spawn_half_size;  /* pick_cube.py:38 */
  float cube_spawn_center[2];
  float cube_half_size;
  float max_goal_height;
  float robot_qpos_noise;      /* table/scene_builder.py:73 robot_init_qpos_noise */
  int32_t n_rest;              /* dofs of articulation 0 */
  float rest_qpos[16];         /* rest configuration; the last two (fingers) are set without noise */
  int32_t obj_fb, goal_fb;     /* free-body indices of the object and the goal site */
} B2SPickReset;

typedef struct B2SPickAutoReset {
  const float* rand;           /* [n_envs, 24] uniform [0,1): 0-1 cube xy, 2 cube yaw, 3-4 goal xy, 5 goal z, 6-23 pairs for 9 Box-Muller normals */
  float* final_obs;            /* [n_envs, obs_dim] observation of the finished episode (rows of sub-scenes with done = 0 are untouched) */
  uint8_t* done;               /* [n_envs] out: 1 where the sub-scene was reset */
  int32_t ignore_terminations; /* 1: only truncation ends an episode (ManiSkillVectorEnv(ignore_terminations=True)) */
  int32_t max_episode_steps;   /* TimeLimit (utils/registration.py:160-168): truncated = elapsed >= max_episode_steps, written to flags[:, 5] */
} B2SPickAutoReset;

int32_t b2s_pick_task_set_reset(uint64_t world, uint64_t handle, const B2SPickReset* reset);
/* the auto-reset alone, after a b2s_pick_task_step (lets the caller render the finished state in between) */
int32_t b2s_pick_task_autoreset(uint64_t world, uint64_t handle, const B2SPickOutputs* out, const B2SPickAutoReset* ar, void* stream);
/* dst[env] = src[env] (row_bytes per sub-scene, a multiple of 16) where mask[env] != 0 -- `final_observation` images of the sub-scenes
 * that are about to be reset (vector/wrappers/gymnasium.py:165 clones the whole observation; here only finished rows move) */
int32_t b2s_masked_copy(uint64_t world, void* dst_dev, const void* src_dev, uint64_t row_bytes, const uint8_t* mask_dev, void* stream);
/* b2s_render for the sub-scenes with env_mask[env] != 0 only (NULL = all): re-render after a partial reset */
int32_t b2s_render_masked(uint64_t world, uint64_t group, const uint8_t* env_mask_dev, void* stream);
/* b2s_pick_task_step followed by the device-side auto-reset; out->flags / reward describe the finished step (what final_info holds) */
int32_t b2s_pick_task_step_autoreset(uint64_t world, uint64_t handle, const float* actions_dev, int32_t substeps, const B2SPickOutputs* out,
                                     const B2SPickAutoReset* ar, void* stream);

/* ---- end-effector controllers: one damped least-squares IK step per sub-scene.
 * Replaces the GPU branch of `Kinematics.compute_ik` (mani_skill/agents/controllers/utils/kinematics.py:197-260: pytorch_kinematics
 * serial-chain Jacobian, then (J^T J + 1e-4 I) dq = J^T delta_pose) for the pd_ee_* control modes (agents/controllers/pd_ee_pose.py:101-133).
 * The chain is the root link -> end link path of the robot: per element the joint origin in the parent link frame, the joint axis in
 * the joint frame and its kind. */
typedef struct B2SChainDesc {
  int32_t n_elem;              /* joints along the chain, fixed ones included, root first */
  const float* origin;         /* [n_elem*7] joint frame in the parent link frame: position xyz + quaternion wxyz */
  const float* axis;           /* [n_elem*3] joint axis in the joint frame */
  const int32_t* kind;         /* [n_elem] 0 fixed, 1 revolute, 2 prismatic */
  const int32_t* qpos_column;  /* [n_elem] column of the joint in the qpos buffer (moving joints; ignored for fixed ones) */
  const uint8_t* controlled;   /* [n_elem] 1 = the joint is solved for (the other moving joints are held; kinematics.py:170-187 qmask) */
  float lambda;                /* damping: 1e-4 in the reference (kinematics.py:245) */
  float alpha;                 /* step scale (solver_config["alpha"]) */
} B2SChainDesc;

int32_t b2s_ik_create(uint64_t world, const B2SChainDesc* chain, uint64_t* ik);
/* delta_pose_dev [n_envs, 6]: translation + XYZ Euler rotation of the end link in the root frame; qpos_dev [n_envs, qpos_stride] the current
 * joint positions; target_dev [n_envs, n_controlled] out: q + alpha dq of the controlled joints, chain order.
 * dq = J^T (J J^T + lambda I)^-1 delta_pose -- the same vector as (J^T J + lambda I)^-1 J^T delta_pose, through the better conditioned
 * 6 x 6 system (Cholesky). */
int32_t b2s_ik_step(uint64_t world, uint64_t ik, const float* delta_pose_dev, const float* qpos_dev, int32_t qpos_stride, float* target_dev,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SIM_H_ */
