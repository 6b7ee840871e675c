/* pk_collide.h — C ABI of the B200-native collision stage for PhysKit.
 *
 * This library replaces, as a drop-in for that path only, the three collision calls of
 * physkit::world::step_impl (reference src/world.cpp:30-46):
 *
 *   broad_phase().update_node(handle, obj.instance().bounds(), vel*dt)     collision_phases.h:371-375
 *   broad_phase().calculate_pairs(narrow_phase(), get_node_handle)         collision_phases.h:377-436
 *   narrow_phase().calculate(get_object, on_beg, on_end)  →  gjk_epa()     collision_phases.h:244-263
 *                                                                           src/collision.cpp:512-518
 *
 * PhysKit has no FFI of its own; its extension point is the pure virtual world_base::step_impl
 * (core/world.h:340).  INTEGRATION.md shows the gpu_world subclass a maintainer would add on top of
 * this header.  Everything is plain pointers and sizes; no C++/torch types cross the boundary.
 *
 * Conventions
 *   - all reals are IEEE double (reference float_t = double, algebra/types.h:8); results are
 *     bit-identical to the reference arithmetic restated in oracle/pk_oracle.hpp (no FMA contraction)
 *   - quaternions are x,y,z,w in memory (Eigen coeffs order, lin_alg.h:388)
 *   - body id = index into the body arrays = the reference's arena slot index (core/world.h:205-206)
 *   - pair key = (min_id << 32) | max_id                       (collision_phases.h:63-69)
 *   - gjk_epa is always called as (a = lower id, b = higher id) (collision_phases.h:252-259)
 *   - every function returns PK_OK (0) or a negative pk_status; nothing throws or aborts
 *   - one ctx = one device + one CUDA stream; a ctx is not re-entrant, distinct ctxs are independent
 *   - there is NO CPU fallback: without a CUDA device pk_create fails with PK_E_NO_DEVICE
 */
#ifndef PK_COLLIDE_H
#define PK_COLLIDE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_ABI_VERSION 1

typedef enum pk_status
{
    PK_OK = 0,
    PK_E_INVALID = -1,       /* bad argument */
    PK_E_NO_DEVICE = -2,     /* no usable CUDA device (never falls back to the CPU) */
    PK_E_CUDA = -3,          /* CUDA runtime error; pk_last_error() has the text */
    PK_E_OOM = -4,           /* device or pinned allocation failed */
    PK_E_PAIR_OVERFLOW = -5, /* candidate pairs exceed pk_config.max_pairs (result.pairs_required set) or GJK hits exceed
                              * max_contacts: pk_reserve_pairs, then repeat the step */
    PK_E_EPA_OVERFLOW = -6,  /* an EPA polytope exceeded the per-pair scratch (result.epa_overflow) */
    PK_E_STATE = -7          /* call order violated (e.g. results requested before pk_collide) */
} pk_status;

/* Pair-set semantics. */
typedef enum pk_mode
{
    /* physkit::world behaviour: persistent fat AABBs (0.1 m margin + displacement, src/bvh.cpp:483-508),
     * a pair exists iff the stored boxes intersect and one member was re-inserted since both exist
     * (collision_phases.h:377-436).  The first step after creation yields no pairs. */
    PK_MODE_WORLD = 0,
    /* dynamic_bvh used directly (tests/dynamic_bvh/main.cpp:600-635): exact boxes, every
     * intersecting (i<j) pair.  Stateless. */
    PK_MODE_QUERY = 1
} pk_mode;

typedef struct pk_config
{
    int32_t device;        /* CUDA device ordinal */
    int32_t mode;          /* pk_mode */
    uint32_t max_bodies;   /* capacity of the body arrays */
    uint32_t max_shapes;   /* capacity of the shape table */
    uint64_t max_pairs;    /* capacity of the candidate-pair list */
    uint64_t max_contacts; /* capacity of the contact list (0 = max_pairs) */
    uint64_t max_hull_vertices; /* total vertices over all hull shapes */
    uint32_t num_worlds;   /* >1: bodies carry a world id, pairs only form inside a world (0/1 = single) */
    uint32_t shard_rank;   /* pair sharding of one world across ranks: this ctx traverses the */
    uint32_t shard_count;  /*   sorted-leaf slice [rank*N/count, (rank+1)*N/count); 0/1 = whole scene */
    uint32_t flags;        /* reserved, 0 */
} pk_config;

typedef struct pk_ctx pk_ctx;
typedef struct pk_multi pk_multi; /* one world sharded over several GPUs from one host process */

/* 88-byte contact record = collision_info (collision.h:52-59) + its pair key.
 * normal points from B to A; depth = EPA face distance (src/collision.cpp:448-453). */
typedef struct pk_contact
{
    uint64_t key;
    double normal[3];
    double world_a[3];
    double world_b[3];
    double depth;
} pk_contact;

/* Closest-distance record of pk_gjk_distance_batch: 64 bytes.  The reference has no such query — gjk_collision
 * (src/collision.cpp:165-189) answers yes / no and gjk_epa returns std::nullopt for a separated pair (:511-518);
 * BASELINE.json's north_star asks for "batched GJK intersection/distance", so this is the distance half, over the same
 * support mappings (bounds.h:164-174, :328-329, :539-548, src/mesh.cpp:341-358, :442-448). */
typedef struct pk_distance
{
    uint64_t key;      /* (a << 32) | b as given */
    double distance;   /* > 0 when the shapes are separated; 0 when they touch or overlap (depth: pk_gjk_epa_batch) */
    double point_a[3]; /* closest points in the world frame, |point_a - point_b| = distance; zeros when not separated */
    double point_b[3];
} pk_distance;

/* contact_point (collision_phases.h:75-88) minus the fields pk_contact already has: the witness points in
 * the bodies' own frames, local = orientation.conjugate() * (world - pos) (core/particle.h:107-108).  This
 * is what narrow_phase::calculate keeps in its manifolds. */
typedef struct pk_contact_point
{
    double local_a[3];
    double local_b[3];
} pk_contact_point;

/* One entry of a world ray cast: what world_base::raycast yields (core/world.h:260-319, a pair of body
 * handle and distance), tagged with the ray of the batch it belongs to. */
typedef struct pk_ray_hit
{
    uint32_t ray;
    uint32_t body;
    double distance; /* ray::intersect_distance of the body's stored box (bvh.h:59-98): 0 when the origin is inside */
} pk_ray_hit;

/* jacobian_row (collision/constraint.h:29-47), the fields constraint_solver::setup_contacts fills.  The normal
 * row's impulse bounds are [0, +inf) (contacts push only, :901), the tangent rows are bounded by the solver's
 * friction cone (:1150-1199). */
typedef struct pk_solver_row
{
    double J_v[3];
    double J_w_a[3];
    double J_w_b[3];
    double M_eff;
    double bias;
} pk_solver_row;

/* contact_solver_point (collision/constraint.h:1204-1214): one per manifold point with positive penetration. */
typedef struct pk_solver_point
{
    uint64_t key;      /* make_pair_key(a, b): a = key >> 32, b = key & 0xffffffff */
    uint32_t manifold; /* index into pk_manifolds() of this step: the reference's `cache` pointer */
    uint32_t point;    /* index of the contact in that manifold */
    pk_solver_row normal, tangent1, tangent2;
    double friction_coeff; /* sqrt(friction_a · friction_b) */
    double inv_m_11, inv_m_12, inv_m_22; /* inverse of the 2×2 friction block, or the diagonal fallback (:1084-1094) */
    double accumulated[3]; /* warm start: normal_impulse, tangent_impulses[0..1] of the manifold point */
} pk_solver_point;

#define PK_RAY_ALL 0     /* every body whose stored box the ray enters within max_distance */
#define PK_RAY_CLOSEST 1 /* per ray the entry with the smallest distance, lowest body id among equals */

/* Contact manifolds: narrow_phase's per-pair state (collision_phases.h:90-327).  A point is
 * manifold::contact_info = contact_point {normal, local_a, local_b, depth} + the solver's cached impulses. */
typedef struct pk_manifold_point
{
    double normal[3];
    double local_a[3];
    double local_b[3];
    double depth;
    double normal_impulse;
    double tangent_impulses[2];
} pk_manifold_point;

typedef struct pk_manifold
{
    uint64_t key;   /* make_pair_key(a, b), a = lower id */
    uint32_t count; /* 1..4 (manifold::max_contact_points); unused points are zero */
    uint32_t _pad;
    pk_manifold_point points[4];
} pk_manifold;

typedef struct pk_manifold_result
{
    uint64_t num_manifolds; /* non-empty manifolds after the update */
    uint64_t num_began;     /* on_collision (empty → non-empty) */
    uint64_t num_ended;     /* on_collision_exit (non-empty → empty while the pair is still in the pair set) */
    float ms;               /* device time of the update */
} pk_manifold_result;

typedef struct pk_step_result
{
    uint64_t num_pairs;      /* candidate pairs produced by the broadphase (this shard) */
    uint64_t num_contacts;   /* pairs for which gjk_epa returned a value */
    uint64_t num_moved;      /* leaves re-inserted this step (|M_moved|, collision_phases.h:374) */
    uint64_t pairs_required; /* on PK_E_PAIR_OVERFLOW: the capacity that would have sufficed */
    uint64_t epa_overflow;   /* pairs whose EPA polytope overflowed the scratch (0 in normal use) */
    uint64_t gjk_hits;       /* pairs whose GJK simplex enclosed the origin (≥ num_contacts) */
    float ms_broadphase;     /* device time of the stages, CUDA events on the ctx stream */
    float ms_narrowphase;
    float ms_total;
    uint32_t step_index;     /* epoch counter of this ctx */
} pk_step_result;

/* Per-kernel device times of the last pk_collide (CUDA events), for bench.py's roofline line. */
#define PK_NUM_STAGES 12
typedef struct pk_stage_times
{
    float ms[PK_NUM_STAGES];
    const char *name[PK_NUM_STAGES];
    uint32_t launches; /* kernels launched by the last pk_collide */
    uint32_t epa_fallback; /* GJK hits epa_coop_kernel handed to epa_kernel (padded simplices, improper horizons) */
} pk_stage_times;

/* ---- lifetime ------------------------------------------------------------------------------ */
int pk_abi_version(void);
int pk_create(const pk_config *cfg, pk_ctx **out);
int pk_destroy(pk_ctx *ctx);
const char *pk_strerror(int status);
const char *pk_last_error(pk_ctx *ctx);

/* ---- shapes (SupportShape models, collision.h:35-38) ------------------------------------------
 * Each returns the new shape id.  Inputs are copied; the caller keeps ownership. */
int pk_shape_box(pk_ctx *ctx, const double half[3], uint32_t *id);       /* obb::support bounds.h:539-548 */
int pk_shape_sphere(pk_ctx *ctx, double radius, uint32_t *id);           /* bounding_sphere::support bounds.h:328-329 */
int pk_shape_hull(pk_ctx *ctx, const double *xyz, uint32_t nverts, uint32_t *id); /* mesh::support mesh.cpp:341-358 */
int pk_shape_aabb(pk_ctx *ctx, const double min[3], const double max[3], uint32_t *id); /* aabb::support bounds.h:164-174; pose-less */
/* Bulk variant: n boxes / spheres in one H2D copy. kind: 1 = box (par = half xyz), 2 = sphere (par[0] = r). */
int pk_shapes_bulk(pk_ctx *ctx, const int32_t *kind, const double *par3, uint32_t n, uint32_t *first_id);

/* Grow the capacities of a live context (never shrinks; 0 for max_contacts = max_pairs).  After PK_E_PAIR_OVERFLOW:
 * pk_reserve_pairs(ctx, result.pairs_required · margin, …) and call pk_collide* again with the same poses — the
 * failed attempt does not advance the context's epoch, and the fat boxes it already updated are left as they are
 * by the second pass, so the repeated step yields the pair set the first would have (only num_moved differs).
 * Results of an earlier step that were not fetched are gone (PK_E_STATE). */
int pk_reserve_pairs(pk_ctx *ctx, uint64_t max_pairs, uint64_t max_contacts);

/* ---- bodies ---------------------------------------------------------------------------------
 * flags bit0 = static (never updated, never queries: src/world.cpp:24), bit1 = alive.
 * A body whose alive bit rises is create_rigid()'d (exact box, not marked moved: core/world.h:202-208,
 * collision_phases.h:342-346); one whose alive bit falls is remove_rigid()'d.
 * disp = vel*dt, the predictive expansion passed to update_node (src/world.cpp:30-31).
 * world_id may be NULL when num_worlds <= 1. */
int pk_bodies_resize(pk_ctx *ctx, uint32_t n);
int pk_bodies_upload(pk_ctx *ctx, const double *pos_xyz, const double *quat_xyzw, const double *disp_xyz,
                     const uint32_t *shape_id, const uint8_t *flags, const uint32_t *world_id,
                     uint32_t first, uint32_t count);
/* Pose-only refresh (the per-step H2D of a running world). Any pointer may be NULL = unchanged.
 * Both calls only ENQUEUE the copies on the context's stream.  Pageable source buffers may be reused when the call
 * returns; page-locked ones (pk_host_alloc) are read by the DMA engine later: leave them untouched until the next
 * call that synchronises the context (pk_collide*, pk_fetch_results, pk_gjk_epa_batch, pk_memcpy_*).
 * With pk_dynamics_enable, a new orientation also refreshes the body's world-frame inertia tensors, as
 * particle::orientation(q) does (core/particle.h:40-44). */
int pk_bodies_update_pose(pk_ctx *ctx, const double *pos_xyz, const double *quat_xyzw, const double *disp_xyz,
                          uint32_t first, uint32_t count);

/* ---- the collision stage --------------------------------------------------------------------
 * pk_collide_resident: run broadphase + narrowphase on the state already in HBM; only the counters
 *   come back to the host.  Results stay on the device.
 * pk_fetch_results:    D2H of pair keys and contact records into ctx-owned pinned memory.
 * pk_collide:          both (what gpu_world::step_impl calls).  The sorted pair keys travel to the host while the
 *                      narrowphase runs and the contact records are stored into the pinned result buffer by the
 *                      EPA kernels themselves, so no copy waits for the end of the step (the buffer holds
 *                      max_contacts records, up to 2 GB; above that, or with PK_NO_MIRROR=1 in the environment,
 *                      the records are copied after the kernels).
 * Pairs are sorted ascending by key; contacts are sorted by key as well.
 * A step reads its counters back once, at its end: the kernels behind the broadphase take the number of candidate pairs
 * from device memory and are launched for 9/8 of the previous step's count (+ 64 Ki).  A step that finds more pairs
 * than that — or a body with more than 64 partners of a larger id, which the per-body pair rows do not hold — is run a
 * second time inside the same call, with the count read back in the middle (and the pair list sorted by a radix sort);
 * results, step_index and num_moved are those of the one step.  Switches for A/B runs, read from the environment at
 * pk_create: PK_SYNC_PAIRS=1 (always read the count back; read per step), PK_PAIR_RADIX=1 (never use pair rows),
 * PK_GJK_EXACT_PREFILTER=1 (round 1's FP64 prefilter instead of the FP32 miss filter), PK_GJK_FILTER_ITERS=n. */
int pk_collide_resident(pk_ctx *ctx, pk_step_result *out);
int pk_fetch_results(pk_ctx *ctx);
int pk_collide(pk_ctx *ctx, pk_step_result *out);
int pk_pairs(pk_ctx *ctx, const uint64_t **keys, uint64_t *n);
int pk_contacts(pk_ctx *ctx, const pk_contact **recs, uint64_t *n);
/* Device-side views for a caller that exchanges results itself (NCCL all-gather of contacts). */
/* Body-local witness points of the contacts of the last step, record k belonging to pk_contacts()[k]
 * (narrow_phase::calculate → contact_point, collision_phases.h:257-263).  Computed on the device on request
 * from the poses of that step; the host pointer stays valid until the next pk_collide*. */
int pk_contact_points(pk_ctx *ctx, const pk_contact_point **pts, uint64_t *n);
/* ---- one world on several GPUs from one host process ------------------------------------
 * pk_create_multi creates one context per listed device (a device may be listed more than once), context i with
 * cfg.shard_rank = i, cfg.shard_count = n: every context holds all bodies and rebuilds the whole tree, context i
 * traverses only its slice of the sorted leaves, so the contexts' pair sets are disjoint and their union is the
 * world's pair set (the sharding bench.py runs with one process per GPU; here the host threads of
 * pk_multi_collide play the ranks).  Shapes are registered on every context through pk_multi_ctx (same calls in
 * the same order give the same ids); the pk_multi_bodies_* calls fan a body upload out to all contexts.
 * pk_multi_collide runs pk_collide on all contexts concurrently, one host thread per context; `total` holds the
 * summed counts and the slowest context's times.  Results are read per context (pk_pairs / pk_contacts on
 * pk_multi_ctx(i)): each is sorted by key, the shards' key ranges interleave.  The first failing context's
 * status is returned. */
int pk_create_multi(const pk_config *cfg, const int *devices, int n, pk_multi **out);
int pk_destroy_multi(pk_multi *m);
int pk_multi_size(pk_multi *m, int *n);
int pk_multi_ctx(pk_multi *m, int i, pk_ctx **ctx);
int pk_multi_bodies_resize(pk_multi *m, uint32_t n);
int pk_multi_bodies_upload(pk_multi *m, const double *pos_xyz, const double *quat_xyzw, const double *disp_xyz,
                           const uint32_t *shape_id, const uint8_t *flags, const uint32_t *world_id, uint32_t first, uint32_t count);
int pk_multi_bodies_update_pose(pk_multi *m, const double *pos_xyz, const double *quat_xyzw, const double *disp_xyz, uint32_t first,
                                uint32_t count);
int pk_multi_collide(pk_multi *m, pk_step_result *total);
/* ---- integrator (optional) ---------------------------------------------------------------
 * The two per-body loops of world::step_impl on the device, so that poses need not travel every step:
 *   loop A (src/world.cpp:22-34): apply_force(gravity · mass), semi_implicit_euler::integrate_vel
 *          (detail/integrate.h:36-40), vel · dt handed to broad_phase::update_node, clear_forces;
 *   loop B (src/world.cpp:50-55): semi_implicit_euler::integrate_pos (detail/integrate.h:42-46).
 * A step of a gpu_world is then  pk_integrate_velocities → pk_collide → [solver] → pk_integrate_positions.
 * pk_dynamics_enable allocates the state (velocities, force accumulators, mass, inertia tensors).
 * pk_dynamics_upload sets it for bodies [first, first+count) whose poses are already uploaded: mass as in
 * particle's constructor (inv_mass = 1 / mass, zero inverse tensor for infinite mass; core/particle.h:14-28),
 * inertia_local = 9 doubles per body, row-major; the inverse is Matrix3d::inverse as the reference computes it.
 * pk_dynamics_set_forces overwrites the accumulators M_acc (= Σ force · inv_mass) and M_torque_acc
 * (particle.h:78-99); pk_dynamics_set_velocities brings a host solver's result back; pk_dynamics_download
 * reads poses and velocities (any pointer may be NULL); pk_displacements reads vel · dt of the last loop A.
 * Static and dead bodies are skipped like in the reference (src/world.cpp:24, 52). */
int pk_dynamics_enable(pk_ctx *ctx);
int pk_dynamics_upload(pk_ctx *ctx, const double *vel, const double *ang_vel, const double *mass, const double *inertia_local,
                       uint32_t first, uint32_t count);
int pk_dynamics_set_velocities(pk_ctx *ctx, const double *vel, const double *ang_vel, uint32_t first, uint32_t count);
int pk_dynamics_set_forces(pk_ctx *ctx, const double *acc, const double *torque, uint32_t first, uint32_t count);
int pk_integrate_velocities(pk_ctx *ctx, double dt, const double gravity[3]);
int pk_integrate_positions(pk_ctx *ctx, double dt);
int pk_dynamics_download(pk_ctx *ctx, double *pos, double *quat, double *vel, double *ang_vel, uint32_t first, uint32_t count);
int pk_displacements(pk_ctx *ctx, double *disp, uint32_t first, uint32_t count);
/* ---- contact rows (optional; needs pk_dynamics_enable and pk_manifolds_enable) -------------
 * constraint_solver::setup_contacts (collision/constraint.h:1052-1104): build_contact_jacobian (:874-953) for
 * every point of every manifold of the last pk_manifolds_update, from the poses and velocities on the device.
 * gravity_norm is the solver's M_gravity (world_desc.gravity().norm(), core/world.h:383): the restitution
 * threshold is 2 · gravity_norm · dt (:1071).  Rows come out ordered by (pair key, point index); points whose
 * penetration is not positive yield no row (:893-894).  The Gauss-Seidel sweep that consumes the rows
 * (:1107-1201) is serial in the reference and stays with the caller.
 * pk_material_upload sets restitution / friction per body (defaults 0.5 / 0.5, core/object.h:107-108). */
int pk_material_upload(pk_ctx *ctx, const double *restitution, const double *friction, uint32_t first, uint32_t count);
int pk_contact_rows_setup(pk_ctx *ctx, double dt, double gravity_norm, uint64_t *nrows);
int pk_contact_rows(pk_ctx *ctx, const pk_solver_point **rows, uint64_t *n);
int pk_contact_rows_device(pk_ctx *ctx, const void **dptr, uint64_t *n, float *device_ms);
/* ---- ray casts --------------------------------------------------------------------------
 * Replaces world_base::raycast(ray, max_dist) (core/world.h:260-319), i.e. dynamic_bvh::raycast (bvh.h:346-450)
 * over the static and the dynamic tree, for a batch of rays against the tree of the LAST step (call after
 * pk_collide / pk_collide_resident; PK_E_STATE otherwise).  origins / directions are nrays xyz triples in host
 * memory (directions need not be normalised: the ray constructor does that, bvh.h:47-50), max_distance one
 * value per ray, world the ray's world in a batched context (NULL: world 0).  The reference yields its
 * entries in an order that depends on the shape of its trees; here they come sorted by (ray, body), with
 * bit-identical distances.  PK_RAY_CLOSEST is the closest-leaf search a caller builds from the callback form
 * (bvh.h:346-398) by returning the entry distance as the new max_distance.
 * hits receives at most `capacity` records; *nhits is the number found.  PK_E_PAIR_OVERFLOW: nothing was
 * written, *nhits is the capacity to retry with.  capacity may not exceed max(max_bodies, max_pairs). */
int pk_raycast(pk_ctx *ctx, const double *origins, const double *directions, const double *max_distance, const uint32_t *world,
               uint32_t nrays, int mode, pk_ray_hit *hits, uint64_t capacity, uint64_t *nhits);
/* Device time of the last pk_raycast (traversal + sort + gather, without the copies), CUDA events. */
int pk_raycast_device_ms(pk_ctx *ctx, float *ms);
/* ---- manifolds (optional) -------------------------------------------------------------
 * narrow_phase::calculate merges each pair's new contact into the manifold it kept from the last step: warm
 * start of a point found at the same place, drift / breaking test of the old points under the new poses,
 * add_reduce beyond four points (collision_phases.h:244-320).  pk_manifolds_enable reserves the state;
 * pk_manifolds_update runs that merge on the device for the step pk_collide* just computed (once per step);
 * manifolds come back sorted by key.  Pairs that left the pair set lose their manifold silently
 * (on_pair_removed, :225-242); began / ended are the keys for which the reference would call on_coll_beg /
 * on_coll_end (:314-318), sorted.  pk_manifolds_set_impulses stores what the constraint solver accumulated
 * (normal, tangent 0, tangent 1 per point; [n][4][3], order of pk_manifolds) for next step's warm start. */
int pk_manifolds_enable(pk_ctx *ctx, uint64_t capacity);
int pk_manifolds_update(pk_ctx *ctx, pk_manifold_result *out);
int pk_manifolds(pk_ctx *ctx, const pk_manifold **recs, uint64_t *n);
int pk_manifolds_device(pk_ctx *ctx, const void **dptr, uint64_t *n);
int pk_manifold_events(pk_ctx *ctx, const uint64_t **began, uint64_t *num_began, const uint64_t **ended, uint64_t *num_ended);
int pk_manifolds_set_impulses(pk_ctx *ctx, const double *impulses, uint64_t n);
int pk_pairs_device(pk_ctx *ctx, const void **dptr, uint64_t *n);
int pk_contacts_device(pk_ctx *ctx, const void **dptr, uint64_t *n);
/* Stored (fat) boxes of the broadphase, [count][6] = min xyz, max xyz (dynamic_bvh::bounds, bvh.h:452-456). */
int pk_stored_bounds(pk_ctx *ctx, double *out6, uint32_t first, uint32_t count);
int pk_stage_times_get(pk_ctx *ctx, pk_stage_times *out);
/* CUDA stream of the ctx as a void* (cudaStream_t) so a host can order its own work after it. */
int pk_stream(pk_ctx *ctx, void **stream);
/* Diagnostic: the library divides the three components of a vector by one norm through a shared, correctly
 * rounded reciprocal (Markstein's division) instead of three IEEE divisions; results must be bit-identical
 * (reference: Eigen normalized(), lin_alg.h:232-240).  Draws `samples` operand pairs on the device — random
 * ones and ones constructed next to exact quotients — and reports how many quotients differ from `/`. */
int pk_selftest_division(pk_ctx *ctx, uint64_t seed, uint64_t samples, uint64_t *mismatches);

/* ---- one world over several PROCESSES, one GPU each -------------------------------------
 * Every rank creates its context with pk_config.shard_rank / shard_count = its rank / the number of ranks and the
 * same capacities, then joins a communicator: rank 0 calls pk_comm_get_id and hands the 128 bytes to the others by
 * whatever means the host has (MPI, a file, torch.distributed), every rank calls pk_comm_init.  The collectives are
 * NCCL all-gathers over NVLink on the context's stream; NCCL is looked up at run time (libnccl.so.2), the library
 * does not link against it.  All ranks must make the same pk_comm_* calls in the same order.
 * Per step: upload the poses of the bodies pk_comm_pose_slice names (1/N of the per-step H2D), pk_comm_allgather_poses,
 * pk_collide_resident, pk_comm_allgather_contacts — the one exchange of the step: every rank then holds the contact
 * records of the whole world, which is what constraint_solver::setup_contacts reads (collision/constraint.h:1052-1104).
 * The single-process form of the same sharding is pk_create_multi below. */
typedef struct pk_comm_id
{
    char bytes[128];
} pk_comm_id;
#define PK_POSE_POS 1
#define PK_POSE_QUAT 2
#define PK_POSE_DISP 4
typedef struct pk_gathered_contacts
{
    const void *d_records;   /* DEVICE pointer: num_ranks blocks of stride_records pk_contact each; block r holds   */
    uint64_t stride_records; /* counts[r] ≤ stride_records records sorted by key (the rest of a block is unspecified),*/
                             /* the ranks' key ranges interleave.  The block size is the same on every rank: the      */
                             /* largest count of the previous exchange plus an eighth, or — first exchange, or a rank */
                             /* outgrew that — the largest count of this one (the records then travel twice)          */
    const uint64_t *counts;  /* host, ctx-owned, valid until the next call                                          */
    uint32_t num_ranks;
    uint64_t total;          /* sum of counts */
    float ms;                /* device time of the exchange (counts + records) */
} pk_gathered_contacts;
int pk_comm_get_id(pk_comm_id *id);
int pk_comm_init(pk_ctx *ctx, const pk_comm_id *id, int rank, int nranks);
int pk_comm_pose_slice(pk_ctx *ctx, uint32_t *first, uint32_t *count); /* bodies [n·r/N, n·(r+1)/N) of this rank */
int pk_comm_allgather_poses(pk_ctx *ctx, int what /* PK_POSE_* bits */); /* in place, enqueued on the ctx stream */
int pk_comm_allgather_contacts(pk_ctx *ctx, pk_gathered_contacts *out);  /* returns when the records have arrived */

/* ---- narrowphase only: gjk_epa over an explicit pair list (BASELINE config C4) ----------------
 * pair_a/pair_b index the uploaded bodies; out[k] / hit[k] are written for every k (host pointers).
 * out[k].key = make_pair_key as given (a<<32|b, NOT min/max: argument order is the caller's). */
int pk_gjk_epa_batch(pk_ctx *ctx, const uint32_t *pair_a, const uint32_t *pair_b, uint64_t n,
                     pk_contact *out, uint8_t *hit);
/* Same with the pair list and outputs resident in HBM (device pointers); returns device ms.
 * Both batch calls share the narrowphase buffers with the step: device-side results of the last pk_collide* that were
 * not fetched can no longer be (pk_fetch_results, pk_contacts_device, pk_manifolds_update, pk_contact_points return
 * PK_E_STATE); host copies already fetched, the pair keys and the tree (pk_raycast) stay valid. */
int pk_gjk_epa_batch_device(pk_ctx *ctx, const uint32_t *d_pair_a, const uint32_t *d_pair_b, uint64_t n,
                            pk_contact *d_out, uint8_t *d_hit, float *ms);

/* ---- narrowphase only: closest distance over an explicit pair list --------------------------------------------
 * The distance form of GJK for every pair (a = pair_a[k], b = pair_b[k]): out[k].distance and the closest points for
 * separated pairs (separated[k] = 1), distance 0 and separated[k] = 0 for pairs that touch or overlap (their depth and
 * normal are pk_gjk_epa_batch's).  Spheres enter as centre + radius, so sphere distances are exact; polytope pairs end
 * on the exact optimum of their vertices' arithmetic (relative gap of the bounds <= 1e-12).  Agrees with
 * gjk_collision's yes / no outside its 1e-6 m margin.  Uses no buffer of the step: results of the last pk_collide*
 * stay valid.  Host pointers; indices >= the number of uploaded bodies return PK_E_INVALID. */
int pk_gjk_distance_batch(pk_ctx *ctx, const uint32_t *pair_a, const uint32_t *pair_b, uint64_t n,
                          pk_distance *out, uint8_t *separated);
/* Same with the pair list and outputs resident in HBM (device pointers); returns device ms.  An index beyond the
 * uploaded bodies yields distance 0, separated 0 for that pair. */
int pk_gjk_distance_batch_device(pk_ctx *ctx, const uint32_t *d_pair_a, const uint32_t *d_pair_b, uint64_t n,
                                 pk_distance *d_out, uint8_t *d_separated, float *ms);

/* Raw device allocation helpers so a host language without CUDA bindings can stage buffers. */
int pk_device_alloc(pk_ctx *ctx, size_t bytes, void **dptr);
int pk_device_free(pk_ctx *ctx, void *dptr);
int pk_memcpy_h2d(pk_ctx *ctx, void *dst, const void *src, size_t bytes);
int pk_memcpy_d2h(pk_ctx *ctx, void *dst, const void *src, size_t bytes);
int pk_memcpy_d2d(pk_ctx *ctx, void *dst, const void *src, size_t bytes); /* e.g. contacts → an NCCL send buffer */
/* Page-locked host memory for the per-step pose arrays (makes pk_bodies_upload a true async DMA). */
int pk_host_alloc(pk_ctx *ctx, size_t bytes, void **hptr);
int pk_host_free(pk_ctx *ctx, void *hptr);

#ifdef __cplusplus
}
#endif
#endif /* PK_COLLIDE_H */
