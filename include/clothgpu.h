/* include/clothgpu.h -- C ABI of the B200-native cloth step + self-collision path.
 *
 * This is the ONLY boundary between host code and the CUDA (sm_100a) kernels: plain pointers and sizes, no C++
 * or torch types.  Every entry point cites the reference interface it replaces (paths relative to the reference
 * repository nobuyuki83/cloth_sim_self_collision).  The source-compatible C++ classes / functions with the
 * reference's own names live in include/clothgpu_compat/ and forward to these symbols.
 *
 * Model: a context owns B independent scenes ("batch") that share one topology (mesh, hinge list, block-CRS
 * pattern, BVH topology) and one rest shape; per-scene state (positions, velocities, matrix values, ILU values,
 * boxes, contact elements, impact zones) is device resident and stays there across steps.  All reals are fp64,
 * all indices int32.  The library never falls back to the CPU: without a CUDA device every compute entry point
 * returns CLOTHGPU_ERR_CUDA.
 *
 * Array conventions (identical to the reference's):
 *   xyz, uvw, xyz0 : np x 3 doubles            tri : ntri x 3 ints             quad : nquad x 4 ints (hinges:
 *                                                                              vertices 2,3 = shared edge)
 *   crs_rowptr[np+1], crs_colidx[ncrs] : block-CRS pattern without the diagonal, columns ascending per row
 *                   (CMatrixSquareSparse::m_colInd / m_rowPtr -- the reference's names are swapped,
 *                    matrix_square_sparse.h:42-44)
 *   valDia : np x 9, valCrs : ncrs x 9 : row-major 3x3 blocks (m_valDia / m_valCrs)
 *   bvh_nodes : nnode x {iroot, ichild0, ichild1} (CNodeBVH, bvh_aabb.h:18-23); leaf: ichild1 == -1 and
 *                   ichild0 == triangle index; node 0 is the root
 *   boxes : nnode x {xmin,xmax,ymin,ymax,zmin,zmax}  (CAABB3D, aabb.h:16-148, without the isnt_empty flag)
 *   elems : n x {ino0,ino1,ino2,ino3,is_fv} in std::set<CContactElement> order (bvh_aabb.h:65-106)
 *   edge_index[np+1], edge_array : vertex->vertex adjacency (CJaggedArray of SetEdgeOfElem(aTri,..),
 *                   self_contact_sparse/main.cpp:574)
 */
#ifndef CLOTHGPU_H
#define CLOTHGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define CLOTHGPU_VERSION 200

typedef struct clothgpu_ctx clothgpu_ctx;

enum {
  CLOTHGPU_OK = 0,
  CLOTHGPU_ERR_ARG = 1,      /* bad argument (null pointer, index out of range, inconsistent sizes) */
  CLOTHGPU_ERR_CUDA = 2,     /* CUDA runtime error or no device; see clothgpu_last_error() */
  CLOTHGPU_ERR_CAPACITY = 3, /* a device work queue overflowed even after growing to its hard limit */
  CLOTHGPU_ERR_STATE = 4     /* call sequence error (e.g. contact ops on a context created without a BVH) */
};

/* colliders shipped with the reference demos (host callbacks there, device built-ins here):
 *   NONE          internal_cloth_sparse/main.cpp:107-110
 *   PLANE_Z       z = par[0]                       self_contact_sparse/main.cpp:110-114
 *   SPHERE        centre par[0..2], radius par[3]  internal_cloth_sparse/main.cpp:120-132
 *   SPHERE_INSIDE same sphere as a container       self_contact_sparse/main.cpp:117-133
 *   TABLE         per-vertex (pd, nx, ny, nz) evaluated by the caller's own callback on the host and uploaded with
 *                 clothgpu_write(CLOTHGPU_BUF_COLLIDER_TABLE) -- the generic path for user callbacks */
enum {
  CLOTHGPU_COLLIDER_NONE = 0,
  CLOTHGPU_COLLIDER_PLANE_Z = 1,
  CLOTHGPU_COLLIDER_SPHERE = 2,
  CLOTHGPU_COLLIDER_SPHERE_INSIDE = 3,
  CLOTHGPU_COLLIDER_TABLE = 4
};

/* scene constants, self_contact_sparse/main.cpp:138-156 and the argument lists of
 * StepTime_InternalDynamicsILU (solve_internal_sparse.h:255-275) / GetIntermidiateVelocityContactResolved
 * (self_collision_cloth.h:21-34) */
typedef struct clothgpu_params {
  double dt;
  double lambda, myu;      /* Lame constants of the membrane (WdWddW_CST) */
  double stiff_bend;       /* WdWddW_Bend */
  double gravity[3];
  double mass_point;
  double stiff_contact;    /* penalty stiffness vs the collider AND cloth_contact_stiffness of the self-collision */
  double contact_clearance;
  int collider_kind;
  int reserved;
  double collider_par[4];
} clothgpu_params;

/* device buffers addressable through clothgpu_read / clothgpu_write (per scene unless noted) */
enum {
  CLOTHGPU_BUF_XYZ = 0,       /* aXYZ  : np*3 doubles */
  CLOTHGPU_BUF_UVW = 1,       /* aUVW  : np*3 doubles */
  CLOTHGPU_BUF_XYZ_PREV = 2,  /* aXYZ1 : positions at the start of the step (self_contact_sparse/main.cpp:224) */
  CLOTHGPU_BUF_VAL_DIA = 3,   /* mat_A.m_valDia : np*9 */
  CLOTHGPU_BUF_VAL_CRS = 4,   /* mat_A.m_valCrs : ncrs*9 */
  CLOTHGPU_BUF_ILU_DIA = 5,   /* ilu_A.mat.m_valDia : np*9 */
  CLOTHGPU_BUF_ILU_CRS = 6,   /* ilu_A.mat.m_valCrs : ncrs*9 */
  CLOTHGPU_BUF_VEC_R = 7,     /* right-hand side / residual : np*3 */
  CLOTHGPU_BUF_VEC_X = 8,     /* solution : np*3 */
  CLOTHGPU_BUF_VEC_P = 9,     /* search direction : np*3 */
  CLOTHGPU_BUF_VEC_Z = 10,    /* preconditioned residual / A*p : np*3 */
  CLOTHGPU_BUF_BOXES = 11,    /* aBB : nnode*6 doubles */
  CLOTHGPU_BUF_ELEMS = 12,    /* contact elements of the last detection pass : 5 ints each */
  CLOTHGPU_BUF_LEAF_PAIRS = 13, /* leaf pairs (BVH node ids) of the last traversal : 2 ints each, unordered */
  CLOTHGPU_BUF_ZONE = 14,     /* rigid-impact-zone label per vertex (-1 or smallest vertex id of the zone) : np ints */
  CLOTHGPU_BUF_UVW0 = 15,     /* aUVWm0 : velocities saved before the impact-zone loop : np*3 */
  CLOTHGPU_BUF_COLLIDER_TABLE = 16, /* CLOTHGPU_COLLIDER_TABLE input : np*4 doubles {pd,nx,ny,nz} */
  CLOTHGPU_BUF_COUNT_ = 17
};

/* per-step statistics, the machine-readable form of the reference's stdout prints
 * (solve_internal_sparse.h:321, self_collision_cloth.cpp:386,415,448) */
enum { CLOTHGPU_NSTATS = 256 };
/* stats[0]=PCG iteration, [1]=#proximity elems, [2]=#CCD impulse passes run, [3]=#RIZ rounds run,
 * [4]=is_impulse_applied, [5]=last "NNode In RIZ", [6]=1 if the RIZ loop ended with "Resolved All Collisions",
 * [7]=diagnostic, not a reference quantity: impulse elements (cumulative over the context's life) whose certified
 *     fast arithmetic had to fall back to the plain IEEE operations (same result either way),
 * [14],[15]=diagnostic (CLOTHGPU_RANK_CHECK=1): narrow-phase hits whose depth-first rank rebuilt from the root paths
 *     differs from the rank the frontier traversal carries (must be 0) / hits checked, cumulative,
 * [8..13)=CCD elems per impulse pass, [16..116)=CCD elems per RIZ round, [128..228)="NNode In RIZ" per RIZ round */

/* ------------------------------------------------------------------------------------------------ life cycle */
const char* clothgpu_last_error(void);
int clothgpu_version(void);
/* which build of the library this is: libclothgpu.so is the parity build (no multiply-add contraction, the
 * reference's serial dot products: bit-compatible); libclothgpu_fast.so is the same source compiled with contraction
 * on and tree reductions -- the non-parity fast mode of SURVEY.md 8(f)-4.  The string starts with "parity" or "fast". */
const char* clothgpu_build_mode(void);
/* number of visible CUDA devices (0 without a driver/GPU); never fails */
int clothgpu_device_count(void);

/* Replaces the init sequence of self_contact_sparse/main.cpp:561-583 (mat_A.Initialize / SetPattern,
 * ilu_A.Initialize_ILU0, aEdge.SetEdgeOfElem, BVH topology) -- everything the step needs that depends on
 * topology only is derived here once: ILU level schedules and update lists, assembly gather lists, BVH refit
 * order and DFS ranks.  bvh_nodes / edge_* may be NULL (nnode = 0) for a context that only runs the implicit
 * step (internal_cloth_sparse).  device < 0 keeps the current device. */
int clothgpu_create(clothgpu_ctx** out, int device, int nscene, int np, int ntri, int nquad, const int* tri,
                    const int* quad, const int* bcflag, const double* xyz0, const int* crs_rowptr,
                    const int* crs_colidx, int nnode, const int* bvh_nodes, const int* edge_index,
                    const int* edge_array);
void clothgpu_destroy(clothgpu_ctx* ctx);

/* CJaggedArray::SetEdgeOfElem + Sort (jagged_array.h:158-207, 71-90), i.e. the input of CMatrixSquareSparse::SetPattern
 * (matrix_square_sparse.cpp:140-163) when called with the hinge quads and aEdge when called with the triangles: the
 * vertex -> vertex adjacency of an element list as CSR with ascending, unique columns, built on the GPU (one 64-bit
 * radix sort + unique of the (row, col) pairs).  elem: nelem x nnoel; index[nnode+1], array[cap] out (array may be
 * NULL to query the size).  Returns the number of entries, or -CLOTHGPU_ERR_*. */
int clothgpu_build_adjacency(int nelem, int nnoel, const int* elem, int nnode, int with_diagonal, int* index,
                             int* array, int cap);

/* MakeBVHTopology_TopDown, bvh_aabb.cpp:294-313 (MakeTriSurTri :109-138, DevideTriAryConnex :156-290): BVH topology
 * over the triangles, node for node the array the reference builds (same splits, same list orders, same numbering)
 * wherever the reference survives; where it crashes (bvh_aabb.cpp:238, exactly square sub-grids) the split falls back
 * to the longest axis with ties resolved x,y,z and *nfallback (may be NULL) counts those nodes.  Host code, no GPU
 * needed.  nodes: cap_nodes x {iroot, ichild0, ichild1}, cap_nodes >= 2*ntri-1.  Returns the node count, or
 * -CLOTHGPU_ERR_* on failure. */
int clothgpu_bvh_build_topdown(int ntri, const int* tri, int np, const double* xyz, int* nodes, int cap_nodes,
                               int* nfallback);

int clothgpu_nscene(const clothgpu_ctx* ctx);
/* element count of a buffer for one scene, in units of its type (doubles or ints) */
long long clothgpu_buf_len(const clothgpu_ctx* ctx, int buf);
/* host <-> device copies of one scene's buffer (synchronous with respect to the context's stream) */
int clothgpu_write(clothgpu_ctx* ctx, int buf, int scene, const void* host, long long count);
int clothgpu_read(clothgpu_ctx* ctx, int buf, int scene, void* host, long long count);
/* aXYZ / aUVW of one scene in one call (the state the reference keeps in two std::vector<double>) */
int clothgpu_upload_state(clothgpu_ctx* ctx, int scene, const double* xyz, const double* uvw);
int clothgpu_download_state(clothgpu_ctx* ctx, int scene, double* xyz, double* uvw);
/* page-locked host memory for the buffers passed to the copies above (optional: pageable memory works, pinned memory
 * lets the copies run as one DMA each -- the reference keeps aXYZ / aUVW in std::vector<double>, main.cpp:36-37) */
int clothgpu_alloc_host(void** ptr, long long bytes);
int clothgpu_free_host(void* ptr);
/* per-scene gravity override (config 5: perturbed copies); scene < 0 clears all overrides */
int clothgpu_set_scene_gravity(clothgpu_ctx* ctx, int scene, const double* gravity);
int clothgpu_synchronize(clothgpu_ctx* ctx);

/* device timers on the context's stream (CUDA events): start/stop bracket a region, elapsed returns ms */
int clothgpu_timer_start(clothgpu_ctx* ctx);
int clothgpu_timer_stop(clothgpu_ctx* ctx, float* ms);
/* number of kernel launches issued by this context so far */
long long clothgpu_launch_count(const clothgpu_ctx* ctx);
/* per-phase device time of the last step call, ms: [0]=assembly [1]=ilu [2]=pcg [3]=refit [4]=traversal+narrow
 * [5]=sort/unique [6]=impulse [7]=riz [8]=velocity fix-up; enabled by clothgpu_set_profiling(ctx, 1) (adds syncs) */
enum { CLOTHGPU_NPHASE = 16 };
int clothgpu_set_profiling(clothgpu_ctx* ctx, int on);
int clothgpu_phase_ms(clothgpu_ctx* ctx, double* ms);
/* per-kernel device time: while on, every kernel launch is bracketed by CUDA events on the context's stream and
 * synchronised (measurement mode, slower).  clothgpu_kernel_times writes one "name launches total_ms" line per kernel
 * into out (NUL terminated). */
int clothgpu_set_kernel_timing(clothgpu_ctx* ctx, int on);
int clothgpu_kernel_times(clothgpu_ctx* ctx, char* out, int cap);

/* ------------------------------------------------------------------------------------------------ implicit step */
/* mat_A.SetZero + AddWdWddW_Cloth + AddWdWddW_Contact + AddWdW_Gravity + rhs + mass + SetBoundaryCondition:
 * solve_internal_sparse.h:281-313.  XYZ, UVW -> VAL_DIA, VAL_CRS, VEC_R */
int clothgpu_assemble(clothgpu_ctx* ctx, const clothgpu_params* p);
/* CMatrixSquareSparse::MatVec, matrix_square_sparse.cpp:167-270: y = alpha*A*x + beta*y on vector buffers */
int clothgpu_matvec(clothgpu_ctx* ctx, double alpha, int buf_x, double beta, int buf_y);
/* InnerProduct, matrix_square_sparse.cpp:313-324; out[nscene].  Fixed-shape tree sum (deterministic run to run;
 * not the reference's serial order) */
int clothgpu_dot(clothgpu_ctx* ctx, int buf_a, int buf_b, double* out);
/* AXPY, matrix_square_sparse.cpp:327-336: y += a*x */
int clothgpu_axpy(clothgpu_ctx* ctx, double a, int buf_x, int buf_y);
/* CPreconditionerILU::SetValueILU, ilu_sparse.cpp:352-384: VAL_* -> ILU_* */
int clothgpu_ilu_setvalue(clothgpu_ctx* ctx);
/* CPreconditionerILU::DoILUDecomp, ilu_sparse.cpp:387-624 (len == 3): in place on ILU_* */
int clothgpu_ilu_factor(clothgpu_ctx* ctx);
/* CPreconditionerILU::Solve, ilu_sparse.h:23-26: forward + backward substitution in place on a vector buffer */
int clothgpu_ilu_solve(clothgpu_ctx* ctx, int buf_vec);
/* Host-only self check of the strip-pipelined triangular-solve plan (csrc/trisolve.cu) that clothgpu_create derives
 * from the block-CRS pattern: builds the forward and backward plans and verifies their invariants.  Needs no GPU.
 * force_ns > 0 fixes the strip count.  stats: NULL or 16 ints, per direction {strips, rows per strip, micro-steps,
 * longest strip, max ghosts, simulated makespan, micro-steps on the pipelined fast path, value bytes per stage} */
int clothgpu_selftest_triplan(int np, const int* crs_rowptr, const int* crs_colidx, int force_ns, int* stats);
/* Self test of the certified fast arithmetic of the impulse chain (csrc/imp_fast.cuh) against the plain IEEE
 * division / square root on `samples` pseudo-random operand pairs (GPU).  out10[2k] = certified results that differ
 * from the hardware result (must be 0), out10[2k+1] = samples that could not be certified (they take the plain
 * operations), for k = 0 quotient from a Newton reciprocal, 1 quotient from the correctly rounded reciprocal,
 * 2 reciprocal, 3 square root, 4 reciprocal of the root. */
int clothgpu_selftest_certified_math(int device, long long samples, unsigned long long seed, long long* out10);
/* The radicand thresholds the impulse chain compares squared lengths against (host code, needs no GPU):
 * *ge = the smallest x with sqrt(x) >= c  (length(v) < c  <=>  v.v < *ge),
 * *le = the largest  x with sqrt(x) <= c  (length(v) > c  <=>  v.v > *le). */
int clothgpu_selftest_sqrt_thresholds(double c, double* ge, double* le);
/* Instrumentation of the triangular-solve kernels: while on, the compute warp of every CTA accumulates cycle counters.
 * out: NULL, or 2 x cap_ctas x 8 values of the last forward / backward launch {start ns, end ns, cycles waiting for
 * matrix chunks, cycles waiting for imported values, compute cycles, signalling cycles, micro-steps, strip};
 * nctas[2] receives the CTA counts.  The call first returns the counters, then sets the new on/off state. */
int clothgpu_debug_trisolve(clothgpu_ctx* ctx, int on, unsigned long long* out, int cap_ctas, int* nctas);
/* Solve_PCG, ilu_sparse.cpp:628-709: VEC_R (destroyed) -> VEC_X.  conv_ratio / iteration: nscene values,
 * in = tolerance / iteration cap (element 0 is used for all scenes), out = achieved ratio / iterations */
int clothgpu_pcg(clothgpu_ctx* ctx, double* conv_ratio, int* iteration);
/* Solve_CG, matrix_square_sparse.cpp:338-406 (unpreconditioned variant, same in/out convention) */
int clothgpu_cg(clothgpu_ctx* ctx, double* conv_ratio, int* iteration);
/* StepTime_InternalDynamicsILU, solve_internal_sparse.h:255-326: one backward-Euler step on XYZ, UVW.
 * conv_ratio / iteration may be NULL; otherwise nscene values out */
int clothgpu_step_internal_ilu(clothgpu_ctx* ctx, const clothgpu_params* p, double* conv_ratio, int* iteration);
/* StepTime_InternalDynamics, solve_internal_sparse.h:185-253 (CG, 1e-4 / 1000) */
int clothgpu_step_internal_cg(clothgpu_ctx* ctx, const clothgpu_params* p, double* conv_ratio, int* iteration);
/* conv_ratio / iteration of the most recent PCG or CG solve of the context (the values the reference prints at
 * solve_internal_sparse.h:248,321), nscene values each; either pointer may be NULL.  No device work. */
int clothgpu_last_solver(clothgpu_ctx* ctx, double* conv_ratio, int* iteration);
/* UpdateIntermidiateVelocity, solve_internal_sparse.h:329-369: XYZ -> UVW */
int clothgpu_update_intermediate_velocity(clothgpu_ctx* ctx, const clothgpu_params* p);

/* ------------------------------------------------------------------------------------------------ element level */
/* The reference's per-element free functions as batches of n independent elements on the GPU (current device, no
 * context, synchronous).  The scalar forms of include/clothgpu_compat/ call them with n = 1; the values are the ones
 * the fused step computes.
 * MakePositiveDefinite_Sim22, cloth_internal_physics.h:15-67: s2, s3 = n x {s00, s11, s01} */
int clothgpu_elem_psd_sym22(int n, const double* s2, double* s3);
/* WdWddW_CST, cloth_internal_physics.h:70-176: C, c = n x [3][3] rest / current positions; W[n]; dW = n x [3][3];
 * ddW = n x [ino][jno][idim][jdim] (81 per element) */
int clothgpu_elem_cst(int n, const double* C, const double* c, double lambda, double myu, double* W, double* dW,
                      double* ddW);
/* WdWddW_Bend, cloth_internal_physics.h:180-238: C, c = n x [4][3] (vertices 2,3 = shared edge); dW = n x [4][3];
 * ddW = n x [4][4][3][3] (144 per element).  W is returned from zero (the reference adds to its argument) */
int clothgpu_elem_bend(int n, const double* C, const double* c, double stiff, double* W, double* dW, double* ddW);
/* WdWddW_Contact, cloth_internal_physics.h:243-280: pd_normal = n x {pd, nx, ny, nz} as returned by the caller's
 * penetrationDepth callback (host code); dW = n x 3; ddW = n x [3][3] */
int clothgpu_elem_contact(int n, const double* pd_normal, double stiff_contact, double contact_clearance, double* W,
                          double* dW, double* ddW);
/* DistanceFaceVertex, bvh_aabb.cpp:321-340: p = n x {p0, p1, p2, p3} (12 doubles) -> dist, w0, w1 */
int clothgpu_dist_face_vertex(int n, const double* p, double* dist, double* w0, double* w1);
/* DistanceEdgeEdge, bvh_aabb.cpp:343-384: p = n x {p0, p1, q0, q1} -> dist, ratio_p, ratio_q */
int clothgpu_dist_edge_edge(int n, const double* p, double* dist, double* ratio_p, double* ratio_q);
/* FindCoplanerInterp, bvh_aabb.cpp:598-662: s, e = n x 4 points at the start / end of the step -> t (or -1) */
int clothgpu_coplanar_interp(int n, const double* s, const double* e, double* t);
/* MakeNormal, self_contact_sparse/main.cpp:183-211: unit vertex normals of one scene from a positions buffer (the
 * demos hand them to the renderer; not part of the timed step).  normals: np*3 doubles out */
int clothgpu_make_normal(clothgpu_ctx* ctx, int buf_xyz, int scene, double* normals);

/* ------------------------------------------------------------------------------------------------ self collision */
/* BuildBoundingBoxChild_Prx, bvh_aabb.cpp:23-61: positions buffer -> BOXES */
int clothgpu_bvh_refit_prx(clothgpu_ctx* ctx, int buf_xyz, double delta);
/* BuildBoundingBoxChild_CCD, bvh_aabb.cpp:64-106 */
int clothgpu_bvh_refit_ccd(clothgpu_ctx* ctx, int buf_xyz, int buf_uvw, double dt);
/* GetContactElement_Proximity, bvh_aabb.cpp:441-557: BOXES, positions -> ELEMS; counts[nscene] out.
 * LEAF_PAIRS holds the leaf pairs the traversal reached. */
int clothgpu_contacts_prx(clothgpu_ctx* ctx, int buf_xyz, double delta, int* counts);
/* GetContactElement_CCD, bvh_aabb.cpp:797-923 */
int clothgpu_contacts_ccd(clothgpu_ctx* ctx, int buf_xyz, int buf_uvw, double dt, double delta, int* counts);
/* number of leaf pairs of the last traversal, per scene */
int clothgpu_leaf_pair_counts(clothgpu_ctx* ctx, int* counts);
/* SelfCollisionImpulse_Proximity, self_collision_cloth.cpp:20-112: ELEMS, positions -> velocities (in place) */
int clothgpu_impulse_prx(clothgpu_ctx* ctx, int buf_xyz, int buf_uvw, double delta, double stiffness, double dt,
                         double mass);
/* SelfCollisionImpulse_CCD, self_collision_cloth.cpp:116-212 */
int clothgpu_impulse_ccd(clothgpu_ctx* ctx, int buf_xyz, int buf_uvw, double delta, double stiffness, double dt,
                         double mass);
/* rigid impact zones: reset = the empty aRIZ of self_collision_cloth.cpp:427; make = MakeRigidImpactZone :220-274
 * on ELEMS; apply = ApplyRigidImpactZone :298-348 (velocities <- rigid motion fitted to UVW0) */
int clothgpu_riz_reset(clothgpu_ctx* ctx);
int clothgpu_riz_make(clothgpu_ctx* ctx);
int clothgpu_riz_apply(clothgpu_ctx* ctx, int buf_xyz, int buf_uvw);
/* GetIntermidiateVelocityContactResolved, self_collision_cloth.cpp:355-456 on XYZ_PREV (positions) and UVW.
 * applied[nscene] out; stats may be NULL, else nscene x CLOTHGPU_NSTATS ints */
int clothgpu_resolve_self_contact(clothgpu_ctx* ctx, const clothgpu_params* p, int* applied, int* stats);

/* ------------------------------------------------------------------------------------------------ whole step */
/* StepTime of self_contact_sparse/main.cpp:213-263 (without MakeNormal), nsteps times, state device resident.
 * stats: NULL or nscene x CLOTHGPU_NSTATS ints of the LAST step */
int clothgpu_step_self_contact(clothgpu_ctx* ctx, const clothgpu_params* p, int nsteps, int* stats);
/* StepTime of internal_cloth_sparse/main.cpp:205-237, nsteps times */
int clothgpu_step_internal(clothgpu_ctx* ctx, const clothgpu_params* p, int nsteps, int* iterations);

#ifdef __cplusplus
}
#endif
#endif /* CLOTHGPU_H */
