/* dex_model_blob.h — layout of the compiled ("flattened") model blob.
 *
 * The reference compiles MJCF to an MjModel through dm_control/MuJoCo every
 * episode (reference dexterity/manipulation/__init__.py:56-86).
 * This build compiles the scene ONCE on the host (dexterity_b200/model/) into a
 * self-describing blob of int32 / float64 arrays which is handed across the
 * C ABI (include/dexterity_b200.h: dex_model_create) and to the CPU oracle.
 *
 * Blob = { header, directory[nfields], payload }.  All offsets are in BYTES from
 * the start of the blob and 8-byte aligned.  Field ids are the positions in
 * DEX_MODEL_FIELDS below; the Python packer parses this very list, so this
 * header is the single source of truth for the layout.
 *
 * Conventions follow MuJoCo: quaternions are (w,x,y,z); joint types
 * FREE=0 BALL=1 SLIDE=2 HINGE=3; geom types PLANE=0 SPHERE=2 CAPSULE=3 BOX=6
 * MESH=7; free joint = 3 world-frame translational dofs then 3 body-frame
 * rotational dofs.
 */
#ifndef DEX_MODEL_BLOB_H_
#define DEX_MODEL_BLOB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEX_BLOB_MAGIC 0x44455842u /* "DEXB" */
#define DEX_BLOB_VERSION 3u

enum { DEX_BLOB_I32 = 0, DEX_BLOB_F64 = 1 };

/* X(name, dtype, doc) — order defines the field id. */
#define DEX_MODEL_FIELDS(X)                                                      \
  X(DIMS, DEX_BLOB_I32, "see DexDim")                                                 \
  X(OPT, DEX_BLOB_F64, "see DexOpt")                                                  \
  X(BODY_PARENT, DEX_BLOB_I32, "nbody")                                               \
  X(BODY_ROOT, DEX_BLOB_I32, "nbody: top-level ancestor (child of world)")            \
  X(BODY_WELD, DEX_BLOB_I32, "nbody: first ancestor-or-self with a joint, else 0")    \
  X(BODY_JNTADR, DEX_BLOB_I32, "nbody (-1 if none)")                                  \
  X(BODY_JNTNUM, DEX_BLOB_I32, "nbody")                                               \
  X(BODY_DOFADR, DEX_BLOB_I32, "nbody (-1 if none)")                                  \
  X(BODY_DOFNUM, DEX_BLOB_I32, "nbody")                                               \
  X(BODY_POS, DEX_BLOB_F64, "nbody*3 frame offset in parent")                         \
  X(BODY_QUAT, DEX_BLOB_F64, "nbody*4")                                               \
  X(BODY_IPOS, DEX_BLOB_F64, "nbody*3 inertial frame in body")                        \
  X(BODY_IQUAT, DEX_BLOB_F64, "nbody*4")                                              \
  X(BODY_MASS, DEX_BLOB_F64, "nbody")                                                 \
  X(BODY_INERTIA, DEX_BLOB_F64, "nbody*3 principal moments")                          \
  X(BODY_INVWEIGHT0, DEX_BLOB_F64, "nbody*2 (trans, rot) at qpos0")                   \
  X(BODY_XFRC, DEX_BLOB_F64, "nbody*6 constant applied (force,torque) at body com")   \
  X(JNT_TYPE, DEX_BLOB_I32, "njnt")                                                   \
  X(JNT_BODY, DEX_BLOB_I32, "njnt")                                                   \
  X(JNT_QPOSADR, DEX_BLOB_I32, "njnt")                                                \
  X(JNT_DOFADR, DEX_BLOB_I32, "njnt")                                                 \
  X(JNT_LIMITED, DEX_BLOB_I32, "njnt")                                                \
  X(JNT_POS, DEX_BLOB_F64, "njnt*3 anchor in body")                                   \
  X(JNT_AXIS, DEX_BLOB_F64, "njnt*3 unit axis in body")                               \
  X(JNT_RANGE, DEX_BLOB_F64, "njnt*2")                                                \
  X(JNT_MARGIN, DEX_BLOB_F64, "njnt")                                                 \
  X(JNT_SOLREF, DEX_BLOB_F64, "njnt*2 limit solref")                                  \
  X(JNT_SOLIMP, DEX_BLOB_F64, "njnt*5 limit solimp")                                  \
  X(QPOS0, DEX_BLOB_F64, "nq")                                                        \
  X(DOF_BODY, DEX_BLOB_I32, "nv")                                                     \
  X(DOF_JNT, DEX_BLOB_I32, "nv")                                                      \
  X(DOF_PARENT, DEX_BLOB_I32, "nv (-1 if none)")                                      \
  X(DOF_ARMATURE, DEX_BLOB_F64, "nv")                                                 \
  X(DOF_DAMPING, DEX_BLOB_F64, "nv")                                                  \
  X(DOF_FRICTIONLOSS, DEX_BLOB_F64, "nv")                                             \
  X(DOF_INVWEIGHT0, DEX_BLOB_F64, "nv")                                               \
  X(DOF_SOLREF, DEX_BLOB_F64, "nv*2 friction-loss solref")                            \
  X(DOF_SOLIMP, DEX_BLOB_F64, "nv*5 friction-loss solimp")                            \
  X(TEN_ADR, DEX_BLOB_I32, "ntendon: first wrap")                                     \
  X(TEN_NUM, DEX_BLOB_I32, "ntendon: number of wraps")                                \
  X(TEN_LIMITED, DEX_BLOB_I32, "ntendon")                                             \
  X(TEN_RANGE, DEX_BLOB_F64, "ntendon*2")                                             \
  X(TEN_MARGIN, DEX_BLOB_F64, "ntendon")                                              \
  X(TEN_INVWEIGHT0, DEX_BLOB_F64, "ntendon")                                          \
  X(TEN_SOLREF, DEX_BLOB_F64, "ntendon*2 limit solref")                               \
  X(TEN_SOLIMP, DEX_BLOB_F64, "ntendon*5 limit solimp")                               \
  X(WRAP_JNT, DEX_BLOB_I32, "nwrap: joint id (fixed tendons only)")                   \
  X(WRAP_COEF, DEX_BLOB_F64, "nwrap")                                                 \
  X(ACT_TRNTYPE, DEX_BLOB_I32, "nu: 0 joint, 1 tendon")                               \
  X(ACT_TRNID, DEX_BLOB_I32, "nu: joint or tendon id")                                \
  X(ACT_CTRLLIMITED, DEX_BLOB_I32, "nu")                                              \
  X(ACT_FORCELIMITED, DEX_BLOB_I32, "nu")                                             \
  X(ACT_GEAR, DEX_BLOB_F64, "nu")                                                     \
  X(ACT_GAINPRM, DEX_BLOB_F64, "nu*3")                                                \
  X(ACT_BIASPRM, DEX_BLOB_F64, "nu*3 affine bias (const, length, velocity)")          \
  X(ACT_CTRLRANGE, DEX_BLOB_F64, "nu*2")                                              \
  X(ACT_FORCERANGE, DEX_BLOB_F64, "nu*2")                                             \
  X(GEOM_TYPE, DEX_BLOB_I32, "ngeom (collidable geoms only)")                         \
  X(GEOM_BODY, DEX_BLOB_I32, "ngeom")                                                 \
  X(GEOM_VERTADR, DEX_BLOB_I32, "ngeom (-1 unless mesh)")                             \
  X(GEOM_VERTNUM, DEX_BLOB_I32, "ngeom")                                              \
  X(GEOM_CONTYPE, DEX_BLOB_I32, "ngeom")                                              \
  X(GEOM_CONAFFINITY, DEX_BLOB_I32, "ngeom")                                          \
  X(GEOM_POS, DEX_BLOB_F64, "ngeom*3 in body frame (mesh: hull centroid)")            \
  X(GEOM_QUAT, DEX_BLOB_F64, "ngeom*4 in body frame")                                 \
  X(GEOM_SIZE, DEX_BLOB_F64, "ngeom*3")                                               \
  X(GEOM_RBOUND, DEX_BLOB_F64, "ngeom bounding-sphere radius about GEOM_POS")         \
  X(VERT, DEX_BLOB_F64, "nvert*3 convex-hull vertices, geom frame")                   \
  X(VERT_ADJADR, DEX_BLOB_I32, "nvert: first hull-graph neighbour in ADJ")            \
  X(VERT_ADJNUM, DEX_BLOB_I32, "nvert")                                               \
  X(ADJ, DEX_BLOB_I32, "nadj: neighbour vertex ids, local to the mesh")               \
  X(PAIR_GEOM1, DEX_BLOB_I32, "npair candidate geom pairs after static filters")      \
  X(PAIR_GEOM2, DEX_BLOB_I32, "npair (type1 <= type2)")                               \
  X(PAIR_CONDIM, DEX_BLOB_I32, "npair")                                               \
  X(PAIR_FRICTION, DEX_BLOB_F64, "npair*5 mixed friction")                            \
  X(PAIR_SOLREF, DEX_BLOB_F64, "npair*2 mixed solref")                                \
  X(PAIR_SOLIMP, DEX_BLOB_F64, "npair*5 mixed solimp")                                \
  X(PAIR_MARGIN, DEX_BLOB_F64, "npair")                                               \
  X(PAIR_GAP, DEX_BLOB_F64, "npair")                                                  \
  X(SITE_BODY, DEX_BLOB_I32, "nsite")                                                 \
  X(SITE_POS, DEX_BLOB_F64, "nsite*3")                                                \
  X(SITE_QUAT, DEX_BLOB_F64, "nsite*4")

enum DexField {
#define DEX_X(name, dtype, doc) DEXF_##name,
  DEX_MODEL_FIELDS(DEX_X)
#undef DEX_X
  DEXF_COUNT
};

/* indices into the DIMS field */
enum DexDim {
  DEXD_NBODY = 0, DEXD_NJNT, DEXD_NQ, DEXD_NV, DEXD_NU, DEXD_NTENDON, DEXD_NWRAP,
  DEXD_NGEOM, DEXD_NVERT, DEXD_NADJ, DEXD_NPAIR, DEXD_NSITE,
  DEXD_NCONMAX, DEXD_SOLVER_ITERATIONS, DEXD_MPR_ITERATIONS, DEXD_LS_ITERATIONS,
  DEXD_COUNT
};

/* indices into the OPT field */
enum DexOpt {
  DEXO_TIMESTEP = 0, DEXO_GRAVITY_X, DEXO_GRAVITY_Y, DEXO_GRAVITY_Z,
  DEXO_TOLERANCE, DEXO_MPR_TOLERANCE, DEXO_IMPRATIO, DEXO_MEANINERTIA,
  DEXO_LS_TOLERANCE,
  DEXO_COUNT
};

typedef struct DexBlobHeader {
  uint32_t magic;
  uint32_t version;
  uint32_t nfields;
  uint32_t reserved;
  uint64_t total_bytes;
} DexBlobHeader;

typedef struct DexBlobEntry {
  uint32_t dtype;   /* DEX_BLOB_I32 or DEX_BLOB_F64 */
  uint32_t reserved;
  uint64_t offset;  /* bytes from blob start */
  uint64_t count;   /* elements */
} DexBlobEntry;

#ifdef __cplusplus
}
#endif
#endif /* DEX_MODEL_BLOB_H_ */
