/*
 * cubezcuda.h — C ABI of libcubezcuda, the B200 (sm_100a) implementation of the cubez
 * per-step rigid-body pipeline.
 *
 * This is the drop-in boundary: the entry points a cgo (or ctypes / JNI) binding of the
 * reference's Go API would bind.  No torch types, no C++ types, no callbacks; plain
 * pointers and sizes.  Every function returns CZ_OK (0) or a negative status code and
 * never throws.  Host pointers passed in are only used for the duration of the call
 * (cgo rule: no Go pointer is retained).  Every entry point does cudaSetDevice(handle's
 * device) because a goroutine may hop OS threads between calls.
 *
 * There is NO CPU fallback behind this ABI: if the CUDA runtime/device is missing the
 * calls fail with CZ_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference root, tbogdala/cubez):
 *   cz_integrate ................ (*RigidBody).Integrate            rigidbody.go:213-259
 *   cz_calculate_derived_data ... (*RigidBody).CalculateDerivedData rigidbody.go:268-299
 *   cz_collider_derive .......... Collision{Cube,Sphere}.CalculateDerivedData
 *                                                                   colliders.go:173-176, 302-304
 *   cz_narrowphase .............. CheckForCollisions + CheckAgainst{HalfSpace,Sphere,Cube}
 *                                                                   colliders.go:180-254, 308-441, 576-747
 *   cz_resolve_contacts ......... ResolveContacts                   contact.go:208-222 (and :59-616)
 *   cz_world_* .................. the per-frame loop of examples/cubedrop.go:29-75 and
 *                                 examples/ballistic.go:27-105, re-expressed as a batched
 *                                 world handle (new API; no reference counterpart)
 *
 * Precision: cz_real is double (libcubezcuda.so) or float (libcubezcuda_f32.so, built with
 * -DCUBEZ_REAL_FLOAT — mirrors editing `type Real float64`, math/math.go:23).
 *
 * Array conventions: host arrays are flat "array of small vectors": position is n*3,
 * orientation n*4 in (w,x,y,z) order (math/math.go:58-59), matrices column-major
 * (math/matrix.go:6-12): Matrix3 n*9, Matrix3x4 n*12.  Any output pointer may be NULL to skip
 * that field.  Body index -1 stands for a nil *RigidBody (plane contacts).
 */
#ifndef CUBEZCUDA_H
#define CUBEZCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef CUBEZ_REAL_FLOAT
typedef float cz_real;
#else
typedef double cz_real;
#endif

/* status codes */
#define CZ_OK 0
#define CZ_ERR_INVALID -1       /* bad handle / argument */
#define CZ_ERR_CUDA -2          /* CUDA runtime error (message in cz_last_error) */
#define CZ_ERR_CAPACITY -3      /* more contacts than the declared capacity (never truncated silently) */
#define CZ_ERR_NIL_BODY -4      /* contact.go:512-523: frictionless one-body contact (Go nil dereference) */
#define CZ_ERR_NOMEM -5

/* collider shape tags (colliders.go:29-71) */
#define CZ_SHAPE_NONE 0
#define CZ_SHAPE_CUBE 1
#define CZ_SHAPE_SPHERE 2

/* pair schedules */
#define CZ_SCHED_ALL_PAIRS_ORDERED 0 /* examples/cubedrop.go:47-64: for i: planes, then every j != i */
#define CZ_SCHED_EXPLICIT 1          /* examples/ballistic.go:47-97: caller-supplied ordered check list */

/* world flags */
#define CZ_WORLD_BROADPHASE 1   /* all-pairs schedule evaluated through the sort-based broadphase */
#define CZ_WORLD_FUSED 2        /* force the fused small-world kernel (error if the world does not fit) */
#define CZ_WORLD_NO_FUSED 4     /* force the multi-kernel path */

/* math-layer op codes for cz_math_op (inputs/outputs are flat cz_real arrays) */
#define CZ_OP_VEC_ADD 1               /* a[3] b[3] -> [3]        math/vector.go:7   */
#define CZ_OP_VEC_ADD_SCALED 2        /* a[3] b[3] s -> [3]      math/vector.go:14  */
#define CZ_OP_VEC_COMPONENT_PRODUCT 3 /* a[3] b[3] -> [3]        math/vector.go:26  */
#define CZ_OP_VEC_CROSS 4             /* a[3] b[3] -> [3]        math/vector.go:33  */
#define CZ_OP_VEC_DOT 5               /* a[3] b[3] -> [1]        math/vector.go:42  */
#define CZ_OP_VEC_MAGNITUDE 6         /* a[3] -> [1]             math/vector.go:47  */
#define CZ_OP_VEC_SQUARE_MAGNITUDE 7  /* a[3] -> [1]             math/vector.go:52  */
#define CZ_OP_VEC_MUL_WITH 8          /* a[3] s -> [3]           math/vector.go:57  */
#define CZ_OP_VEC_NORMALIZE 9         /* a[3] -> [3]             math/vector.go:64  */
#define CZ_OP_VEC_SUB 10              /* a[3] b[3] -> [3]        math/vector.go:82  */
#define CZ_OP_QUAT_MUL 20             /* q[4] p[4] -> [4]        math/quaternion.go:46 */
#define CZ_OP_QUAT_LEN 21             /* q[4] -> [1]             math/quaternion.go:41 */
#define CZ_OP_QUAT_NORMALIZE 22       /* q[4] -> [4]             math/quaternion.go:78 */
#define CZ_OP_QUAT_ROTATE 23          /* q[4] v[3] -> [3]        math/quaternion.go:56 */
#define CZ_OP_QUAT_ADD_SCALED_VECTOR 24 /* q[4] v[3] s -> [4]    math/quaternion.go:20 */
#define CZ_OP_M3_MUL_M3 30            /* a[9] b[9] -> [9]        math/matrix.go:89  */
#define CZ_OP_M3_INVERT 31            /* a[9] -> [9]             math/matrix.go:133 */
#define CZ_OP_M3_MUL_V 32             /* a[9] v[3] -> [3]        math/matrix.go:80  */
#define CZ_OP_M3_TRANSFORM_TRANSPOSE 33 /* a[9] v[3] -> [3]      math/matrix.go:157 */
#define CZ_OP_M3_DETERMINANT 34       /* a[9] -> [1]             math/matrix.go:127 */
#define CZ_OP_M34_MUL_M34 35          /* a[12] b[12] -> [12]     math/matrix.go:198 */
#define CZ_OP_M34_MUL_V 36            /* a[12] v[3] -> [3]       math/matrix.go:188 */
#define CZ_OP_M34_TRANSFORM_INVERSE 37 /* a[12] v[3] -> [3]      math/matrix.go:222 */
#define CZ_OP_M34_SET_AS_TRANSFORM 38 /* pos[3] q[4] -> [12]     math/matrix.go:167 */
#define CZ_OP_REAL_EQUAL 40           /* a b -> [1] (0/1)        math/math.go:64    */
#define CZ_OP_TRANSFORM_INERTIA 41    /* iitBody[9] rot[12] -> [9] rigidbody.go:275 */

typedef struct cz_ctx cz_ctx;
typedef struct cz_world cz_world;

/* RigidBody state, rigidbody.go:23-101.  mass itself is host-only (GetMass). */
typedef struct cz_bodies {
    int32_t n;
    cz_real *position;                     /* n*3 */
    cz_real *orientation;                  /* n*4 */
    cz_real *velocity;                     /* n*3 */
    cz_real *rotation;                     /* n*3 */
    cz_real *acceleration;                 /* n*3 */
    cz_real *linear_damping;               /* n   */
    cz_real *angular_damping;              /* n   */
    cz_real *inverse_inertia_tensor;       /* n*9, body space */
    cz_real *inverse_mass;                 /* n   */
    cz_real *motion;                       /* n   */
    uint8_t *is_awake;                     /* n   */
    uint8_t *can_sleep;                    /* n   */
    cz_real *transform;                    /* n*12, derived */
    cz_real *inverse_inertia_tensor_world; /* n*9,  derived */
    cz_real *last_frame_acceleration;      /* n*3 */
} cz_bodies;

/* CollisionCube / CollisionSphere, colliders.go:39-71 */
typedef struct cz_colliders {
    int32_t n;
    int32_t *shape;     /* n, CZ_SHAPE_* */
    int32_t *body;      /* n, index into the accompanying cz_bodies (object API); world API: collider i <-> body i */
    cz_real *offset;    /* n*12 */
    cz_real *transform; /* n*12, derived */
    cz_real *half_size; /* n*3 (cubes) */
    cz_real *radius;    /* n   (spheres) */
} cz_colliders;

/* CollisionPlane, colliders.go:29-35 */
typedef struct cz_planes {
    int32_t n;
    cz_real *normal; /* n*3 */
    cz_real *offset; /* n   */
} cz_planes;

/* Contact, contact.go:17-51 (public fields).  `n` is in/out: capacity on input arrays is
 * `capacity`, number of valid contacts is `n`. */
typedef struct cz_contacts {
    int32_t capacity;
    int32_t n;
    int32_t *body0;       /* -1 = nil */
    int32_t *body1;       /* -1 = nil */
    cz_real *friction;
    cz_real *restitution;
    cz_real *point;       /* n*3 */
    cz_real *normal;      /* n*3 */
    cz_real *penetration; /* n */
    int32_t *check;       /* optional: index of the check that produced the contact */
} cz_contacts;

typedef struct cz_world_desc {
    int32_t n_worlds;
    int32_t bodies_per_world;   /* body slots per world (unused slots: shape NONE / active_from = INT32_MAX) */
    int32_t contacts_per_world; /* contact capacity per world; exceeding it is CZ_ERR_CAPACITY */
    int32_t schedule;           /* CZ_SCHED_* */
    int32_t flags;              /* CZ_WORLD_* */
} cz_world_desc;

typedef struct cz_step_stats {
    int64_t steps;           /* world-steps executed by the call (n_worlds * n_steps) */
    int64_t contacts;        /* contacts generated, summed */
    int64_t pos_iterations;  /* adjustPositions iterations used, summed (contact.go:233) */
    int64_t vel_iterations;  /* adjustVelocities iterations used, summed (contact.go:390) */
    int64_t checks;          /* narrowphase checks evaluated, summed */
    int64_t kernel_launches; /* kernels launched by the call */
    int32_t max_contacts;    /* largest per-world contact count seen */
    int32_t status;          /* CZ_OK or the first error raised on the device */
    float device_ms;         /* CUDA-event time of the call's kernels on the world's stream */
} cz_step_stats;

/* ---- context ---------------------------------------------------------------------- */
int cz_init(int device, cz_ctx **out);
int cz_shutdown(cz_ctx *ctx);
const char *cz_last_error(cz_ctx *ctx); /* ctx may be NULL: last error of the calling thread */
int cz_real_size(void);                 /* sizeof(cz_real) of this build: 8 or 4 */
void *cz_ctx_stream(cz_ctx *ctx);       /* the cudaStream_t the context launches on */
int cz_ctx_synchronize(cz_ctx *ctx);
int cz_host_alloc(cz_ctx *ctx, uint64_t bytes, void **out); /* pinned host memory for zero-copy views */
int cz_host_free(cz_ctx *ctx, void *p);

/* ---- object-API shims: host buffers in, host buffers out (batch of n, n may be 1) -- */

/* RigidBody.Integrate for io->n bodies (rigidbody.go:213-259).  lin_pow/ang_pow/bias are the
 * three math.Pow results of :233,:234,:250, evaluated by the HOST LANGUAGE so they are
 * bit-identical to what it would have computed itself; pass NULL / NaN to let the library
 * evaluate them with C pow() in float64. */
int cz_integrate(cz_ctx *ctx, cz_bodies *io, cz_real dt, const cz_real *lin_pow, const cz_real *ang_pow,
                 const cz_real *bias);
/* RigidBody.CalculateDerivedData (rigidbody.go:268-299) */
int cz_calculate_derived_data(cz_ctx *ctx, cz_bodies *io);
/* collider.transform = body.transform x Offset (colliders.go:173-176, 302-304) */
int cz_collider_derive(cz_ctx *ctx, int32_t n, const cz_real *body_transform, const cz_real *offset,
                       cz_real *out_transform);
/* CheckForCollisions over an ordered list of checks (colliders.go:720-747).  one[k]/two[k]:
 * >= 0 collider index, < 0 plane -(p+1).  Contacts are appended to `out` in check order then
 * vertex order (the order the reference's append calls produce); found[k] (optional) is the
 * bool each check returns.  bodies supplies Velocity for the cube-sphere fallback normal
 * (colliders.go:417-421). */
int cz_narrowphase(cz_ctx *ctx, const cz_colliders *colliders, const cz_planes *planes, const cz_bodies *bodies,
                   int32_t n_checks, const int32_t *one, const int32_t *two, cz_contacts *out, uint8_t *found);
/* ResolveContacts(maxIterations, contacts, duration) (contact.go:208-222).  Mutates bodies
 * (Position, Orientation, Velocity, Rotation, IsAwake, motion, and transform + world inertia
 * for bodies that were asleep) and contacts (Penetration; Bodies swap + negated normal when
 * body0 was nil).  iterations_used[2] (optional) = {position, velocity} iterations. */
int cz_resolve_contacts(cz_ctx *ctx, int32_t max_iterations, cz_contacts *io, cz_bodies *bodies_io, cz_real dt,
                        int32_t *iterations_used);

/* ---- batched-world handle ------------------------------------------------------------ */
int cz_world_create(cz_ctx *ctx, const cz_world_desc *desc, cz_world **out);
int cz_world_destroy(cz_world *w);
/* Upload bodies for worlds [first_world, first_world+n_worlds): arrays hold
 * n_worlds*bodies_per_world entries, world-major.  Derived fields (transform, world inertia,
 * last_frame_acceleration) are uploaded when given; pass derive != 0 to have the device run
 * CalculateDerivedData on them instead. */
int cz_world_upload_bodies(cz_world *w, int32_t first_world, int32_t n_worlds, const cz_bodies *b, int32_t derive);
int cz_world_upload_colliders(cz_world *w, int32_t first_world, int32_t n_worlds, const cz_colliders *c,
                              int32_t derive);
int cz_world_upload_planes(cz_world *w, const cz_planes *p);                      /* shared by all worlds */
/* Explicit ordered check list: one[k]/two[k] >= 0 body (collider) index, < 0 plane -(p+1).  Planes must be uploaded
 * first: an entry naming a plane that does not exist is CZ_ERR_INVALID. */
int cz_world_upload_schedule(cz_world *w, int32_t n_checks, const int32_t *one, const int32_t *two);
/* active_from[i]: body i takes part from this step index on; integrate[i] == 0: body is never
 * integrated (ballistic backboard, examples/ballistic.go:27-44).  NULL = all 0 / all 1. */
int cz_world_set_activation(cz_world *w, int32_t first_world, int32_t n_worlds, const int32_t *active_from,
                            const uint8_t *integrate);
/* Override the Pow factors (rigidbody.go:233,234,250) for duration dt with host-language values. */
int cz_world_set_pow(cz_world *w, cz_real dt, const cz_real *lin_pow, const cz_real *ang_pow, cz_real bias);
int cz_world_set_step_index(cz_world *w, int64_t step_index);
/* Per-pair surface materials (new API).  The reference hard-wires Friction = 0.9 and
 * Restitution = 0.1 into every generated contact ("test constants", the FIXME at colliders.go:199-202 and
 * :246-249, :358-361, :435-438, :509-512, :702-705); this replaces the constants with a table lookup.  Every
 * collider and every plane carries a material id in [0, n_materials); the contacts produced by
 * CheckForCollisions(one, two) take Friction = friction[m(one) * n_materials + m(two)] and Restitution likewise
 * (operands as the schedule names them; a symmetric table makes the order irrelevant).  body_material holds
 * n_worlds * bodies_per_world ids for worlds [first_world, first_world + n_worlds) (NULL: unchanged),
 * plane_material one id per uploaded plane (NULL: unchanged; upload the planes first).  A new n_materials resets
 * every id to 0.  n_materials = 0 restores the constants.  Friction 0 selects calculateFrictionlessImpulse with
 * the reference's defect mirrored (contact.go:498-531: CZ_ERR_NIL_BODY for a one-body contact). */
int cz_world_set_materials(cz_world *w, int32_t n_materials, const cz_real *friction, const cz_real *restitution,
                           int32_t first_world, int32_t n_worlds, const int32_t *body_material, const int32_t *plane_material);
/* Force / torque input (new API).  RigidBody keeps forceAccum / torqueAccum (rigidbody.go:86-92), reads
 * them in Integrate (:219-223: lastFrameAcceleration = Acceleration + forceAccum * inverseMass, angular acceleration =
 * inverseInertiaTensorWorld * torqueAccum) and clears them (:206), but nothing in the reference ever writes them.  This
 * is the writer: forceAccum += force[3i..], torqueAccum += torque[3i..] for every body of worlds [first_world,
 * first_world + n_worlds) (host arrays of n_worlds * bodies_per_world * 3 reals; either may be NULL).  The next
 * Integrate of an AWAKE body consumes and clears its accumulators; a sleeping body keeps them (Integrate returns before
 * ClearAccumulators, :214-216).  Until the first call the accumulators do not exist and cost nothing. */
int cz_world_add_forces(cz_world *w, int32_t first_world, int32_t n_worlds, const cz_real *force, const cz_real *torque);
/* RL-style episodes (new API): snapshot the current device state as the episode start; world k
 * is at frame phase0[k] (0 <= phase0[k] < length) of its episode now and is restored to the
 * snapshot at the start of every frame on which its phase wraps to 0.  A reset also clears the force / torque
 * accumulators of the world (forces still waiting — of sleeping bodies, or added just before that frame — do not
 * survive it).  length <= 0 disables. */
int cz_world_set_episodes(cz_world *w, int32_t length, const int32_t *phase0);
/* n_steps frames of updateCallback (examples/cubedrop.go:69-75). Asynchronous unless stats != NULL.
 * Device-side errors (CZ_ERR_CAPACITY, CZ_ERR_NIL_BODY) are STICKY: an asynchronous step cannot return them, so the
 * first one raised stays recorded until a call observes it — a step with stats != NULL, cz_world_synchronize, any
 * cz_world_download_*, cz_world_last_step_counts or cz_world_checksum_energy — which returns it and clears it.
 * A world that overflowed its contact capacity skipped ResolveContacts for that frame: treat the error as fatal for
 * the run and re-create the world with a larger contacts_per_world. */
int cz_world_step(cz_world *w, cz_real dt, int32_t n_steps, cz_step_stats *stats);
int cz_world_synchronize(cz_world *w); /* waits for the world's stream; returns (and clears) a pending device-side error */
int cz_world_download_bodies(cz_world *w, int32_t first_world, int32_t n_worlds, cz_bodies *out);
int cz_world_download_colliders(cz_world *w, int32_t first_world, int32_t n_worlds, cz_colliders *out);
/* contacts of the last step of one world, as generated (before ResolveContacts), canonical order */
int cz_world_download_contacts(cz_world *w, int32_t world, cz_contacts *out);
/* Renderer-side export: the per-body copy the example loop makes each frame,
 * SetGlVector3(&Node.Location, &body.Position) and SetGlQuat(&Node.LocalRotation, &body.Orientation)
 * (examples/cubedrop.go:35-37, examples/exampleapp.go:146-159), done on the device: float32(x) of every
 * component.  location: n*3, rotation: n*4 as (W, V[0], V[1], V[2]), model (optional extra): the body transform
 * as a column-major 4x4; n = n_worlds * bodies_per_world; any pointer may be NULL.  dst_on_device != 0: the
 * pointers are device pointers (a mapped GL buffer) written in place, asynchronously on the context stream. */
int cz_world_export_gl(cz_world *w, int32_t first_world, int32_t n_worlds, float *location, float *rotation, float *model,
                       int32_t dst_on_device);
/* per-world counters of the last step (each array n_worlds long, any may be NULL) */
int cz_world_last_step_counts(cz_world *w, int32_t *n_contacts, int32_t *pos_iterations, int32_t *vel_iterations);
/* NaN / overflow watch: the number of bodies whose position, orientation, velocity or rotation holds a non-finite
 * component.  The reference propagates NaN silently (e.g. 0/0 when both bodies of a contact are static); so does the
 * library, bit for bit — this is how a host notices. */
int cz_world_count_nonfinite(cz_world *w, int64_t *bodies);
/* Contact-island statistics of a CZ_WORLD_BROADPHASE world since creation: frames whose ResolveContacts ran as one CTA
 * per island, and how many of those had to be re-run on the single-CTA path because the reference's loop would have
 * been cut by its iteration cap (the only case islands cannot reproduce).  Either pointer may be NULL. */
int cz_world_island_stats(cz_world *w, int64_t *island_frames, int64_t *fallbacks);
/* FNV-1a-64 of each world's state summed mod 2^64, and total energy. */
int cz_world_checksum_energy(cz_world *w, uint64_t *checksum, double *energy);
/* Host-buffer step: upload primary state, run n_steps, download state, all on the world's
 * stream with pinned staging (the end-to-end call a host-resident caller makes). */
int cz_world_step_host(cz_world *w, cz_bodies *io, cz_real dt, int32_t n_steps, cz_step_stats *stats);
/* RL-style step for device-resident worlds: actions in, observations out.
 * Before the first of the n_steps frames every body i receives Velocity.Add(add_velocity[3i..])
 * and Rotation.Add(add_rotation[3i..]) — the batched form of RigidBody.AddVelocity / AddRotation
 * (rigidbody.go:195-202); either array may be NULL (no call).  After the frames, every non-NULL
 * array among obs->{position, orientation, velocity, rotation, motion, is_awake, transform,
 * inverse_inertia_tensor_world, last_frame_acceleration} is filled (obs->n = n_worlds *
 * bodies_per_world; obs may be NULL).  Host arrays should be pinned; transfers are pipelined in
 * chunks against the step.  Episode resets (cz_world_set_episodes) happen on the device. */
int cz_world_step_rl(cz_world *w, const cz_real *add_velocity, const cz_real *add_rotation, cz_bodies *obs, cz_real dt,
                     int32_t n_steps, cz_step_stats *stats);

/* Pipelined RL loop (new API): the same step, asynchronous.  cz_world_step_rl_async enqueues the actions' upload, the
 * frames and the observations' download and returns a ticket at once; cz_world_rl_wait(ticket) blocks until the
 * observation arrays of that step are complete.  Up to TWO steps may be in flight (wait on the older ticket first):
 * while the observations of step t travel to the host, the frames of step t+1 already run — the device staging buffers
 * alternate, so use two sets of (pinned) host arrays as well and do not touch a set whose step is in flight.  No other
 * call on the world while steps are in flight.  obs32 (optional): observations as float32, converted on the device —
 * position 3, orientation 4 (w,x,y,z), velocity 3, rotation 3 per body; half the bytes of the Real arrays.  The stats
 * of cz_world_rl_wait are cumulative since the previous wait with stats (frames of the younger step may be included). */
typedef struct cz_obs32 {
    int32_t n;
    float *position, *orientation, *velocity, *rotation; /* any may be NULL */
} cz_obs32;
int cz_world_step_rl_async(cz_world *w, const cz_real *add_velocity, const cz_real *add_rotation, cz_bodies *obs, cz_obs32 *obs32, cz_real dt,
                           int32_t n_steps, int32_t *ticket);
int cz_world_rl_wait(cz_world *w, int32_t ticket, cz_step_stats *stats);

/* ---- multi-GPU runs of batched worlds (new API) -----------------------------------------------
 * Worlds share no state, so a batch is partitioned by world index into contiguous shards, one per device: shard k
 * of n owns worlds [k*W/n, (k+1)*W/n).  ONE host process drives every shard (each on its own stream of its own
 * device); stepping exchanges nothing between devices.  cz_run_finish computes the checksum / energy per shard on
 * its device and reduces {u64 checksum, f64 energy, i64 counters, f32 device time} over the shards with ONE grouped
 * ncclAllReduce on a single-process communicator (ncclCommInitAll; libnccl.so.2 is loaded at cz_run_create, the
 * library does not link against it).  Shards that share a device (n_shards devices not all distinct — useful on a
 * one-GPU box) are reduced on the host instead and used_nccl reports 0. */
typedef struct cz_run cz_run;
typedef struct cz_run_totals {
    uint64_t checksum;       /* FNV-1a-64 of every world's state, summed mod 2^64: identical for any shard count */
    double energy;           /* total energy (summation order differs with the shard count: compare to 1e-9 relative) */
    int64_t world_steps;     /* world-steps executed since the run was created or last finished */
    int64_t contacts, pos_iterations, vel_iterations;
    float max_device_ms;     /* slowest shard: CUDA events on its stream, first step enqueued -> last step done */
    int32_t n_shards;
    int32_t used_nccl;       /* 1: reduced by ncclAllReduce; 0: one shard, or shards sharing a device (host reduce) */
} cz_run_totals;
/* devices: n_shards device indices (NULL: 0 .. n_shards-1); desc->n_worlds is the TOTAL over all shards */
int cz_run_create(int32_t n_shards, const int32_t *devices, const cz_world_desc *desc, cz_run **out);
int cz_run_destroy(cz_run *r);
/* shard k: its world handle (for any cz_world_* call on that shard) and the world range it owns */
int cz_run_shard(cz_run *r, int32_t k, cz_world **world, int32_t *first_world, int32_t *n_worlds);
/* whole-batch uploads (arrays hold every world of the run, world-major): sliced per shard */
int cz_run_upload_bodies(cz_run *r, const cz_bodies *all, int32_t derive);
int cz_run_upload_colliders(cz_run *r, const cz_colliders *all, int32_t derive);
int cz_run_upload_planes(cz_run *r, const cz_planes *p);
int cz_run_set_episodes(cz_run *r, int32_t length, const int32_t *phase0);
/* n_steps frames on every shard; launches are enqueued on all devices before any is waited for (asynchronous) */
int cz_run_step(cz_run *r, cz_real dt, int32_t n_steps);
/* wait for every shard, reduce, report; resets the counters and the timer */
int cz_run_finish(cz_run *r, cz_run_totals *out);
const char *cz_run_last_error(cz_run *r);

/* ---- microbench / diagnostics -------------------------------------------------------- */
/* Integrate + CalculateDerivedData over n device-resident free bodies, `steps` times, timed
 * with CUDA events on the context stream; returns average ms per step. */
int cz_bench_integrate(cz_ctx *ctx, int64_t n, uint64_t seed, int32_t warmup, int32_t steps, cz_real dt,
                       float *avg_ms, uint64_t *checksum);
/* Measured FP64 pipe rate of the device: thread-level DMUL / DADD instructions per second with every SM busy and no
 * memory traffic — the denominator of the fused world step's roofline (bench.py), which is FP64-issue bound. */
int cz_bench_fp64_rate(cz_ctx *ctx, double *ops_per_s);
/* Sort-based broadphase (K2) on host-supplied bounding spheres: candidate pairs (i, j) whose
 * spheres overlap within the library's 0.5 % inflation.  pairs holds 2*capacity ints; *n_pairs is
 * the number found (may exceed capacity: CZ_ERR_CAPACITY).  Order is unspecified. */
int cz_broadphase_pairs(cz_ctx *ctx, int64_t n, const cz_real *centers, const cz_real *radii, int64_t capacity,
                        int32_t *pairs, int64_t *n_pairs);
/* K2 microbench: n unit spheres scattered uniformly at `fill` volume fraction (splitmix64 seed),
 * device resident; times bounds -> keys -> radix sort -> cell ranges -> neighbour sweep with CUDA
 * events; returns average ms per frame and the number of candidate pairs. */
int cz_bench_broadphase(cz_ctx *ctx, int64_t n, uint64_t seed, double fill, int32_t warmup, int32_t steps, float *avg_ms,
                        int64_t *n_pairs, float *sort_ms);
/* hand-written radix sort of (key, value) pairs on host buffers (test hook for cz_sort.cuh) */
int cz_sort_pairs_u32(cz_ctx *ctx, int64_t n, uint32_t *keys, uint32_t *vals);
int cz_sort_pairs_u64(cz_ctx *ctx, int64_t n, uint64_t *keys, uint32_t *vals, int32_t bits);

/* device math self-test: runs op `op` of the math layer on one thread (tests port the
 * reference's math/ *_test.go known answers through this). */
int cz_math_op(cz_ctx *ctx, int32_t op, const cz_real *in, cz_real *out);

#ifdef __cplusplus
}
#endif
#endif /* CUBEZCUDA_H */
