/* dexterity_b200.h — C ABI of the B200-native batched physics backend (libdexb200.so).
 *
 * The reference has no FFI layer of its own: its boundary to the physics engine is the
 * `dm_control.mjcf.Physics` object handed to every composer hook plus module-level
 * `mujoco.mj_*` calls.  Each entry point below names the reference call
 * site(s) it replaces.  Conventions:
 *   - every function returns 0 on success, a negative DEX_E_* code on failure; the message is
 *     available from dex_last_error() (thread-local).  No C++ exception crosses the ABI.
 *   - all tensors are caller-owned device memory (PyTorch storage), bound once with dex_bind
 *     as raw pointers + row stride (in elements); the library never allocates per-env state.
 *   - launches are asynchronous on the caller's CUDA stream (`stream` = cudaStream_t, 0 for
 *     the legacy default stream); there is no hidden synchronisation.
 *   - per-env physics failures never raise: they set bits in the bound STATUS tensor
 *     (mirrors dm_control's PhysicsError → episode termination).
 *   - there is NO CPU execution path: without a CUDA device every call fails with DEX_E_CUDA.
 */
#ifndef DEXTERITY_B200_H_
#define DEXTERITY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct DexModel DexModel;
typedef struct DexBatch DexBatch;

enum DexError {
  DEX_OK = 0,
  DEX_E_ARG = -1,       /* bad argument / API misuse */
  DEX_E_MODEL = -2,     /* model blob rejected (unsupported feature, capacity exceeded) */
  DEX_E_CUDA = -3,      /* CUDA runtime error */
  DEX_E_UNBOUND = -4    /* a tensor required by the call has not been bound */
};

/* Bits of the bound STATUS tensor (sticky until the env is reset).  BADQACC and PLACEMENT end the episode with reward 0 and
 * discount 0 on the next post-step (dm_control's non-raising PhysicsError handling); the others are diagnostics. */
enum DexStatusBit {
  DEX_STATUS_BIT_BADQACC = 1,         /* mj_checkAcc: non-finite or huge qacc                                  */
  DEX_STATUS_BIT_CON_OVERFLOW = 2,    /* more than dex_capacity(DEX_CAP_MAXCON) contacts: the rest were dropped */
  DEX_STATUS_BIT_CAND_OVERFLOW = 4,   /* broadphase candidate / staging buffers full                           */
  DEX_STATUS_BIT_SOLVER_MAXITER = 8,  /* Newton stopped at the model's iteration cap                           */
  DEX_STATUS_BIT_CHOL_FAIL = 16,      /* non-positive pivot in the Hessian factorisation (clamped)             */
  DEX_STATUS_BIT_PLACEMENT = 32       /* no contact-free prop pose within DexTaskParams.place_max_attempts     */
};

enum DexDtype { DEX_F32 = 0, DEX_F64 = 1 }; /* F64 = validation build of the same kernels */

/* Bindable tensors.  Shapes are [num_envs][width] row-major with the given row stride. */
enum DexFieldId {
  DEX_FIELD_QPOS = 0,          /* real [nenv][nq]   physics.bind(joints).qpos, data.qpos           */
  DEX_FIELD_QVEL = 1,          /* real [nenv][nv]   data.qvel                                       */
  DEX_FIELD_CTRL = 2,          /* real [nenv][nu]   physics.bind(actuators).ctrl (mujoco_actuation.py:33) */
  DEX_FIELD_QACC_WARMSTART = 3,/* real [nenv][nv]   data.qacc_warmstart                             */
  DEX_FIELD_TIME = 4,          /* real [nenv]       physics.data.time (task.py:149,161,184)         */
  DEX_FIELD_STATUS = 5,        /* u32  [nenv]       DEX_STATUS_* bits                               */
  DEX_FIELD_CONTACT = 6,       /* real [nenv][DEX_MAXCON][8] dist,pos3,normal3,pair  (data.contact) */
  DEX_FIELD_NCON = 7,          /* i32  [nenv]       data.ncon                                       */
  DEX_FIELD_NITER = 8,         /* i32  [nenv]       solver iterations of the last step              */
  DEX_FIELD_DBG_M = 9,         /* real [nenv][nv*nv]      optional: data.qM (dense)                 */
  DEX_FIELD_DBG_QFRC_SMOOTH = 10, /* real [nenv][nv]      optional                                  */
  DEX_FIELD_DBG_QACC = 11,     /* real [nenv][nv]         optional: data.qacc                       */
  DEX_FIELD_DBG_QFRC_CONSTRAINT = 12, /* real [nenv][nv]  optional                                  */
  DEX_FIELD_DBG_EFC = 13,      /* real [nenv][DEX_MAXCON][8] optional: per-contact D, aref[4], mu, imp, R */
  DEX_FIELD_COUNT = 40
};

/* Compiled-capacity constants (also exported through dex_capacity for ctypes users). */
enum DexCapacity { DEX_CAP_MAXCON = 0, DEX_CAP_CON_STRIDE = 1, DEX_CAP_MAXDOF = 2, DEX_CAP_MAXBODY = 3 };
int dex_capacity(int which);

/* Model: replaces mj_loadXML/mj_compile reached through composer.Environment.__init__
 * (reference dexterity/manipulation/__init__.py:81).  `blob` is the host blob of
 * include/dex_model_blob.h; it is copied to `device`. */
int dex_model_create(const void* blob, uint64_t nbytes, int device, DexModel** out);
void dex_model_destroy(DexModel* model);

/* Batch of independent environments (the leading env dimension; replaces MjData). */
int dex_batch_create(DexModel* model, int num_envs, int dtype, DexBatch** out);
void dex_batch_destroy(DexBatch* batch);
int dex_bind(DexBatch* batch, int field, void* dev_ptr, int64_t row_stride);

/* nsub x mj_step on every env: replaces the `n_sub_steps x physics.step()` loop inside
 * composer.Environment.step (reached from dexterity/environment.py:25-29) and
 * physics.step() in goals/fingertip_position.py:97. */
int dex_mj_step(DexBatch* batch, int nsub, void* stream);

/* mj_forward without integration: refreshes contacts (+ optional DBG tensors).  Replaces
 * physics.forward() (goals/fingertip_position.py:89,110; models/hands/dexterous_hand.py:165)
 * and the trailing mj_step1 of dm_control's legacy step that post-step hooks read
 * (utils/mujoco_collisions.py:106 iterates data.contact). */
int dex_mj_forward(DexBatch* batch, void* stream);

/* Collision only (position stage): contacts + ncon. */
int dex_mj_collide(DexBatch* batch, void* stream);

/* Cuts the env range into `nchunk` (1..4) contiguous chunks that run as independent kernel pipelines on
 * library-owned streams, forked from and joined to the caller's stream inside every call (no host
 * synchronisation; results are identical for every nchunk).  Default 1. */
int dex_set_chunks(DexBatch* batch, int nchunk);

/* Chooses the form of the tree kernel K0: 0 (default) = eight lanes per env when the model's tree fits
 * its schedule, 1 = always one thread per env.  Both compute the same quantities. */
int dex_set_tree_kernel(DexBatch* batch, int thread_per_env);

/* Chooses the constraint-solver kernels.  0 = the dense warp-per-env solver for every env; 1 (default) = envs without
 * contacts take the structured eight-lanes-per-env solver (contact-free form), the rest the dense one; 2 = in addition, envs
 * with at most 8 contacts of which at most 2 join two fingers take the structured solver's full form; 3 = envs whose contacts
 * leave the Hessian in the mass matrix's chained no-fill pattern (no contact between two fingers) take the sparse L^T D L form
 * of the warp-per-env kernel, the rest the dense one (measured slower than mode 1, kept as a cross-check).  All minimise the
 * same cost to the same tolerance. */
int dex_set_solver(DexBatch* batch, int mode);

/* Number of this library's kernels launched since batch creation (bench `gpu_launches`). */
int64_t dex_launch_count(const DexBatch* batch);


/* ---------------------------------------------------------------------------------------------
 * Task layer: the reference's composer hooks fused into pre-/post-step kernels.
 * ------------------------------------------------------------------------------------------- */
// Plain-data task description (mirrors the module-level constants of the reference tasks).
typedef struct DexTaskParams {
  int task;                       // DEX_TASK_REORIENT = 0, DEX_TASK_REACH = 1
  int successes_needed;           // reorient.py:64 / reach.py:62
  int steps_before_changing_goal; // reorient.py:70
  int fall_termination;           // reorient.py:103
  int has_max_time;               // max_time_per_goal is not None
  int prop_body, prop_qposadr, prop_dofadr, prop_geom, ground_geom;
  int nhand;                      // number of hand joints (qpos/qvel 0..nhand-1)
  int nsite;                      // fingertip sites (<= 5)
  int site_id[5];
  int dense_reward;               // reach only
  int place_max_attempts;         // reorient: collision-checked prop placement at episode start (dm_control
                                  // initializers.PropPlacer, reorient.py:143-151): poses rejected while the prop is in
                                  // contact, up to this many draws (PropPlacer default 20); 0 = not checked (the caller
                                  // has proven the spawn region contact-free, dexterity_b200/placement.py)
  double success_threshold;       // reorient.py:52 / reach.py:50
  double max_time_per_goal;       // reorient.py:68
  double time_limit;              // manipulation/__init__.py:83 (inf if None)
  double spawn_lo[3], spawn_hi[3];// reorient.py:72-78
  double w_orientation, w_success_bonus, w_action_smoothing, orientation_eps, bonus_threshold;  // reorient.py:49-56
  double reach_tanh_weight;       // reach, dense reward: w of rewards.tanh_squared(d, margin=0.1, loss_at_margin=0.95) =
                                  // atanh(sqrt(loss_at_margin)) / margin (shared/rewards.py:18-28, reach.py:196-205)
  unsigned long long seed;
} DexTaskParams;


/* Additional bindable tensors for the task layer (DexFieldId continues). */
enum DexTaskFieldId {
  DEX_FIELD_OBS = 14,          /* real [nenv][128]   observation vector (layout: dexterity_b200/envs.py)       */
  DEX_FIELD_REWARD = 15,       /* real [nenv]        Task.get_reward                                            */
  DEX_FIELD_DISCOUNT = 16,     /* real [nenv]        Task.get_discount                                          */
  DEX_FIELD_GOAL = 18,         /* real [nenv][16]    GoalTask._goal                                             */
  DEX_FIELD_TASK_I = 19,       /* i32  [nenv][8]     successes, change counter, registered, exceeded, goal_changed, failure, step_type, episode_len */
  DEX_FIELD_TASK_F = 20,       /* real [nenv][8]     solve_start_time, episode_return, goal_distance[<=5]       */
  DEX_FIELD_RNG = 21,          /* u32  [nenv][4]     Philox key0,key1,counter,-                                 */
  DEX_FIELD_ACTION = 22,       /* real [nenv][nu]    the action passed to Environment.step                      */
  DEX_FIELD_EPISODE_STATS = 24,/* f64  [8]           sum return, sum length, episodes, successes, falls, timeouts, physics errors */
  DEX_FIELD_RESET_MASK = 25,   /* u8   [nenv]        optional: envs to reset on the next dex_env_step; the pre-step
                                  consumes the request (clears the byte) when it resets the env                       */
  DEX_FIELD_SKIP = 26,         /* u8   [nenv]        scratch: envs reset during the current call                */
  DEX_FIELD_JC = 27,           /* real [nenv][3*DEX_MAXCON][32] scratch: contact-frame Jacobians (stays in L2)       */
  DEX_FIELD_MSCR = 28,         /* real [nenv][32][32] scratch: dense mass matrix; MUST be zero-initialised by the caller (only
                                  the structural non-zeros are rewritten each step)                                    */
  DEX_FIELD_XKIN = 29,         /* real [nenv][DEX_MAXBODY][16] scratch: body poses xpos(3) xquat(4) xmat(9) (data.xpos/xquat/xmat) */
  DEX_FIELD_TREE = 30,         /* real [nenv][8][32] scratch: cdof (rows 0-5), qfrc_smooth (row 6), tree reference points (row 7) */
  DEX_FIELD_STAGE = 31,        /* real [nenv][64][8] scratch: contact staging slots of the collision kernels               */
  DEX_FIELD_NSLOT = 17,        /* i32  [nenv]        scratch: staging slots in use                                        */
  DEX_FIELD_ORDER = 32,        /* i32  [8][nenv]     optional scratch: env lists per solver-cost class (K2 processes the heaviest first)  */
  DEX_FIELD_TIME64 = 33,       /* f64  [nenv][2]     physics.data.time accumulated in double exactly as MuJoCo does (time += timestep
                                  once per mj_step), and the task's _solve_start_time (task.py:149,161): the per-goal timeout and
                                  the time limit compare THESE, so episode lengths do not depend on the batch dtype.  Required by
                                  the env entry points, optional for dex_mj_step (without it TIME accumulates in the batch dtype);
                                  TIME always receives the rounded value                                                          */
  DEX_FIELD_WORK = 23          /* i32  [16 + 2*cap]  scratch: narrowphase work list: item count, fetch cursor, class counters, then (env, slot<<24|pair)
                                  pairs; bind with row_stride = cap (the capacity in items, >= nenv * 8 recommended)     */
};

/* Installs the task description (copied).  Replaces the task constructors
 * (reference manipulation/tasks/reorient.py:94-168, reach.py:74-139). */
int dex_task_configure(DexBatch* batch, const DexTaskParams* params);

/* Resets every env and writes the FIRST TimeStep (obs, reward=0, discount=1, step_type=FIRST):
 * GoalEnvironment.reset (reference dexterity/environment.py:14-23). */
int dex_env_reset(DexBatch* batch, void* stream);

/* One control step on every env: pre-step (auto-reset of envs whose last TimeStep was LAST, goal
 * switch, action -> ctrl), nsub x mj_step, position/velocity refresh + post-step (reward,
 * discount, termination, observations): GoalEnvironment.step (reference environment.py:25-34). */
int dex_env_step(DexBatch* batch, int nsub, void* stream);

/* The same step split into phases so that a host-side GoalGenerator can act between them (the reach
 * task samples goals with a masked physics sub-rollout, reference goals/fingertip_position.py:61-119):
 * phases = DEX_PHASE_PRE | DEX_PHASE_PHYSICS | DEX_PHASE_POST (any non-empty subset, run in that order). */
enum DexPhase { DEX_PHASE_PRE = 1, DEX_PHASE_PHYSICS = 2, DEX_PHASE_POST = 4 };
int dex_env_phase(DexBatch* batch, int phases, int nsub, int force_reset, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched inverse kinematics (reference inverse_kinematics/ik_solver.py:71-228,
 * controllers/dls/dls.py:43-77): damped-least-squares position IK for all fingertips jointly.
 * ------------------------------------------------------------------------------------------- */
typedef struct DexIkParams {
  int nsite;                 /* number of end-effector sites (<= 5)                                  */
  int site_id[5];            /* hand.fingertip_sites                                                 */
  int max_steps;             /* ik_solver.py:75 (default 100)                                        */
  int num_attempts;          /* ik_solver.py:77 (default 30)                                         */
  int early_stop;            /* ik_solver.py:76                                                      */
  int stop_on_first;         /* ik_solver.py:78                                                      */
  double linear_tol;         /* ik_solver.py:74 (default 1e-3)                                       */
  double gain;               /* _LINEAR_VELOCITY_GAIN = 0.95 (ik_solver.py:20-23)                    */
  double regularization;     /* _REGULARIZATION_WEIGHT = 1e-5 (ik_solver.py:27)                      */
  double progress_threshold; /* _PROGRESS_THRESHOLD = 20 (ik_solver.py:31)                           */
  unsigned long long seed;   /* restart sampling (the reference draws from the global np.random)     */
} DexIkParams;

/* Solves `n` independent target sets.  targets_dev: real [n][nsite*3] (world frame of the IK model);
 * qpos_out_dev: real [n][row_stride] receives the joint configuration (mid-range where no solution);
 * ok_dev: u8 [n] (0 = IKSolver.solve would return None); attempts_dev: optional i32 [n].
 * Replaces IKSolver.solve (ik_solver.py:71-153). */
int dex_ik_solve(DexModel* model, int dtype, const void* targets_dev, int n, const DexIkParams* params,
                 void* qpos_out_dev, int64_t row_stride, unsigned char* ok_dev, int* attempts_dev, void* stream);

/* Per-kernel-class device timing for the benchmark's roofline block: when enabled every launch is
 * bracketed by a cudaEvent pair on the launching stream.  dex_get_kernel_times synchronises and
 * returns total milliseconds and launch counts in arrays of 10: {pre-step, broadphase, dense solve, post-step, tree,
 * narrowphase, contact compaction, structured solve (full form), prop placement, structured solve (contact-free form)}. */
int dex_set_profiling(DexBatch* batch, int on);
int dex_get_kernel_times(DexBatch* batch, double* ms10, int64_t* count10, int reset);

const char* dex_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* DEXTERITY_B200_H_ */
