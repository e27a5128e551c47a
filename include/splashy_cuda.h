/* splashy_cuda.h -- C ABI of libsplashy_cuda.so
 *
 * B200 (sm_100a) implementation of the per-timestep hot path of
 * cantsin/splashy: Convection.apply -> Pressure.calculate -> Pressure.apply
 * (Simulator.fs:82,88-89), plus the validators of Simulator.fs:91-97.
 *
 * The reference has no FFI of its own; these entry points are
 * what its F# host binds with [<DllImport("splashy_cuda")>] in place of the
 * bodies of the functions cited on each declaration (see INTEGRATION.md for
 * the P/Invoke shim).  Plain C: opaque context, plain pointers and sizes, no
 * CUDA / torch types.  There is no CPU fallback: every entry point needs a
 * CUDA device and fails with SPL_E_CUDA otherwise.
 *
 * Layout contract:
 *   - dense box nx * ny * nz, x fastest:  idx = (k * ny + j) * nx + i
 *   - u[idx] = x-velocity on the +x face of cell (i,j,k) (between i and i+1),
 *     v, w likewise (Pressure.fs:18-31)
 *   - media: 0 Air, 1 Fluid, 2 Solid (Cell.fs:9), 255 = cell absent from the
 *     sparse grid; absent / out-of-box behaves like Solid and is never counted
 *     as a neighbour (Pressure.fs:34-38)
 *   - pressure p is defined on Fluid cells; Air is 0 (Constants.fs:20)
 *   - dtype selects fp32 or fp64 for every field pointer of a context
 *
 * A context is not thread-safe; calls are synchronous (they return after the
 * work is complete), like the reference's single-threaded blocking model.
 */
#ifndef SPLASHY_CUDA_H
#define SPLASHY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPL_ABI_VERSION 1

/* ---- enums ------------------------------------------------------------ */
enum { SPL_F32 = 0, SPL_F64 = 1 };                          /* spl_desc.dtype  */
enum { SPL_PC_NONE = 0, SPL_PC_JACOBI = 1, SPL_PC_RBIC0 = 2,
       SPL_PC_MG = 3 /* geometric multigrid V-cycle (opt-in) */ }; /* .precond */
enum { SPL_AIR = 0, SPL_FLUID = 1, SPL_SOLID = 2, SPL_ABSENT = 255 };

/* quirk mask: reproduce the reference's literal Convection.fs behaviour
 * (SURVEY.md Q3-Q5).  0 = intended semi-Lagrangian semantics.              */
enum {
  SPL_Q_REF_INTERP    = 1u << 0, /* Convection.fs:35,42-47 offsets + int-snap key */
  SPL_Q_FORWARD_TRACE = 1u << 1, /* Convection.fs:54,56 sign                       */
  SPL_Q_NEAREST_COPY  = 1u << 2, /* Convection.fs:57,63-64 whole-vector copy       */
  /* marker stage, literal readings of Simulator.fs / Grid.fs / Build.fs */
  SPL_Q_MARKER_SUM    = 1u << 3, /* Simulator.fs:24-35 marker velocity = SUM of the two
                                    faces (intended: their mean)                    */
  SPL_Q_MOVE_CELLS    = 1u << 4, /* Grid.fs:42-47 the cell record travels with its
                                    marker, the old coordinate is deleted           */
  SPL_Q_LAZY_BUFFER   = 1u << 5  /* Build.fs:83-86 air buffer grows from cells that
                                    existed before the call only (one ring per step) */
};

/* return codes; text from spl_last_error mirrors the reference's failwith   */
enum {
  SPL_OK              =  0,
  SPL_E_BADARG        = -1,
  SPL_E_CUDA          = -2,  /* CUDA / NCCL runtime error, or no device             */
  SPL_E_TRACE         = -3,  /* Convection.fs:65 "Backwards particle trace went too far." */
  SPL_E_NONFINITE     = -4,  /* Vector.fs:17 / Pressure.fs:138-142                  */
  SPL_E_DIVERGENCE    = -5,  /* Pressure.fs:149 "Velocity field is not correct"     */
  SPL_E_NOT_CONVERGED = -6,  /* PCG hit max_iters (no reference counterpart)        */
  SPL_E_STATE         = -7,  /* call order (e.g. project before upload)             */
  SPL_E_MARKER        = -8   /* Build.fs:162-174 "Fluid marker went outside bounds" /
                                "does not have a neighbor"; marker left the dense box */
};

/* ---- descriptor (replaces the [<Literal>]s of Constants.fs:10-36) -------- */
typedef struct spl_desc {
  int      nx, ny, nz;     /* this context's box (one z-slab when world > 1)  */
  double   h;              /* Constants.fs:11  cell size                      */
  double   rho;            /* Constants.fs:15  fluid_density                  */
  int      dtype;          /* SPL_F32 | SPL_F64                               */
  int      precond;        /* SPL_PC_*                                        */
  double   tol;            /* stop when max|r| <= tol * max|b|; 0 = never     */
  int      max_iters;      /* PCG iteration cap                               */
  double   tau;            /* SPL_PC_RBIC0: relaxed-MIC weight in [0,1)       */
  unsigned quirks;         /* SPL_Q_* mask                                    */
  int      origin[3];      /* world cell index of box cell (0,0,0); used by
                              SPL_Q_REF_INTERP only (Coord.fs:49-61)          */
  int      device;         /* CUDA device ordinal                             */
  /* z-slab decomposition, one process per GPU (world = 1: single GPU)       */
  int      rank, world;
  int      nz_global;      /* total planes over all slabs (0 => nz)           */
  int      z0;             /* global index of this slab's first plane         */
  const void* nccl_id;     /* world > 1: 128-byte ncclUniqueId from
                              spl_nccl_unique_id, identical on every rank     */
  double   gravity[3];     /* Constants.fs:26, used by spl_add_forces         */
  double   viscosity;      /* Constants.fs:13 fluid_viscosity (kinematic)     */
} spl_desc;

/* result block (SURVEY.md section 5 "metrics")                              */
typedef struct spl_info {
  int      iters;          /* PCG iterations run                              */
  int      converged;      /* 1 if the stopping test was met                  */
  double   r_inf;          /* max |r| at exit                                 */
  double   b_inf;          /* max |b|                                         */
  double   max_abs_div;    /* max |in - out| over Fluid after projection      */
  int64_t  nonfinite;      /* non-finite u,v,w,p entries                      */
  int64_t  trace_escapes;  /* departure points that left the halo / box       */
  float    ms_advect, ms_rhs, ms_solve, ms_grad, ms_total; /* CUDA events     */
  int64_t  kernel_launches;/* kernels launched by the last call               */
} spl_info;

typedef struct spl_ctx spl_ctx;

/* ---- lifetime ----------------------------------------------------------- */
void spl_desc_default(spl_desc* d);  /* Constants.fs values, fp32, Jacobi     */
int  spl_create(spl_ctx** out, const spl_desc* desc);
void spl_destroy(spl_ctx* ctx);
/* message of the last failure on ctx (or of the last failed spl_create when
 * ctx == NULL); the text mirrors the reference's failwith strings
 * (Convection.fs:65, Pressure.fs:138-149, Vector.fs:17).                    */
const char* spl_last_error(const spl_ctx* ctx);
int  spl_abi_version(void);
int  spl_device_count(void);
/* fills 128 bytes; rank 0 calls it and hands the bytes to every rank         */
int  spl_nccl_unique_id(void* out128);

/* ---- state: replaces Grid.fs:12 (the global Dictionary<Coord,Cell>) ------- */
/* host -> device.  media is required; u, v, w, p may be NULL (= zeros).
 * Replaces Grid.add_cells / update_velocities / update_pressures
 * (Grid.fs:38,49-59) as the way state enters the solver.                    */
int  spl_upload(spl_ctx* ctx, const uint8_t* media, const void* u,
                const void* v, const void* w, const void* p);
/* device -> host; every pointer may be NULL.  div = (incoming - outgoing) of
 * Pressure.fs:16-32 for the current velocities (0 on non-Fluid cells).      */
int  spl_download(spl_ctx* ctx, void* u, void* v, void* w, void* p, void* div);
/* The same two transfers without waiting for them (PINNED host buffers): the copies are
 * queued on per-context copy streams, one per direction, and the call returns.
 * spl_upload_async: the next entry point that needs the state (spl_step, ...) waits for
 * the copies on the device and then builds the label-derived data; the host buffers must
 * stay untouched until then.  spl_download_async: spl_sync waits for it.  The copies of
 * one context overlap the kernels of another: with two boxes in flight per GPU
 *     upload_async(B); step(A); download_async(A); upload_async(A'); step(B); ...
 * the PCIe time of Grid <-> device (Grid.fs:38,49-59 in, update_velocities /
 * update_pressures out) hides behind the other box's step.  Every collective of a slab
 * context stays on its compute stream, in program order.                              */
int  spl_upload_async(spl_ctx* ctx, const uint8_t* media, const void* u,
                      const void* v, const void* w, const void* p);
int  spl_download_async(spl_ctx* ctx, void* u, void* v, void* w, void* p);
int  spl_sync(spl_ctx* ctx);

/* Cell-centred scalar fields carried by the flow (north_star: "semi-Lagrangian advection of
 * the staggered MAC velocity and scalar fields"; SURVEY 8d cfg2 "one scalar phi", +8 B/cell).
 * The reference's Cell record holds none (Cell.fs:11-16), so there is no function to replace:
 * a scalar takes the same RK2 backtrace as a velocity face (Convection.fs:50-57), started at the
 * cell centre, and a trilinear sample of itself at the departure point (Convection.fs:21-39;
 * absent corner = 0); like the velocity, only Fluid cells are updated.  fields = n host
 * arrays of the box (context dtype; a NULL entry = zeros), n <= SPL_MAX_SCALARS; every later
 * spl_advect / spl_step / spl_advance advects them with the velocity of the step.          */
#define SPL_MAX_SCALARS 8
int  spl_set_scalars(spl_ctx* ctx, int n, const void* const* fields);
int  spl_get_scalars(spl_ctx* ctx, int n, void* const* fields);

/* ---- the hot path -------------------------------------------------------- */
/* Convection.apply + Grid.update_velocities (Convection.fs:60-66,
 * Simulator.fs:82)                                                           */
int  spl_advect(spl_ctx* ctx, double dt, spl_info* info);
/* Forces.apply (Forces.fs:6-12, Simulator.fs:84): v += gravity * dt on Fluid */
int  spl_add_forces(spl_ctx* ctx, double dt);
/* Viscosity.apply (Viscosity.fs:12-36, Simulator.fs:86) as a Jacobi update;
 * quirk Q10 of SURVEY.md kept (no 1/h^2, always -6 v, sign-gated neighbours)  */
int  spl_add_viscosity(spl_ctx* ctx, double dt);
/* Build.cleanup (Simulator.fs:100-110): propagate velocities into the 2-layer
 * buffer (Build.fs:105-141, order-free reading -- DESIGN.md), zero components
 * pointing into Solid cells (Build.fs:144-160), zero Solid velocities          */
int  spl_cleanup(spl_ctx* ctx);
/* the solver part of Simulator.advance (Simulator.fs:82-110): advect, forces,
 * viscosity, project (+ checks with the reference's 1e-4 threshold when
 * `sanity` != 0), cleanup                                                      */
int  spl_advance(spl_ctx* ctx, double dt, int sanity, spl_info* info);
/* (when markers are set, spl_advance starts with spl_move_markers + spl_relabel:
 * the whole of Simulator.advance, Simulator.fs:54-110)                         */
/* Pressure.divergence over every Fluid cell (Pressure.fs:16-32) -> kept on the
 * device as the PCG right-hand side b = (h rho/dt) * div (Pressure.fs:64-68) */
int  spl_divergence(spl_ctx* ctx, double dt, spl_info* info);
/* Pressure.calculate + Grid.update_pressures + Pressure.apply +
 * Grid.update_velocities (Pressure.fs:41-72,115-129; Simulator.fs:88-89)     */
int  spl_project(spl_ctx* ctx, double dt, spl_info* info);
/* Pressure.check_pressures + check_divergence (Pressure.fs:132-149);
 * returns SPL_E_NONFINITE / SPL_E_DIVERGENCE like the reference throws.      */
int  spl_check(spl_ctx* ctx, double div_tol, spl_info* info);
/* advect -> rhs -> PCG -> gradient in one call (Simulator.fs:82,88-89)        */
int  spl_step(spl_ctx* ctx, double dt, spl_info* info);

/* ---- benchmarking helpers ------------------------------------------------ */
/* page-locked host buffers for the upload / download calls                   */
void* spl_host_alloc(size_t bytes);
void  spl_host_free(void* p);
/* ---- marker stage: the step before the path ---------------------
 * Simulator.fs:19-52,60-79.  Markers are fp64 locations in metres held on the device;
 * cell of a location = Coord.construct(float) (Coord.fs:49-67) minus desc.origin.
 * z-slabs (world > 1): every rank passes and holds ALL markers; a marker is moved by the
 * rank whose slab holds its cell and one all-reduce hands every rank the new locations;
 * bounds (spl_set_world) are box indices of the GLOBAL box.  SPL_Q_MOVE_CELLS: world = 1. */
/* World.contains (World.fs:11-14, Aabb.fs:13-17) as inclusive box-index bounds; cells
 * created outside become Solid (Build.fs:95-98).  Default: the whole box.          */
int  spl_set_world(spl_ctx* ctx, const int lo[3], const int hi[3]);
/* Simulator.generate's state: n marker locations, xyz = n x 3 doubles (metres)     */
int  spl_set_markers(spl_ctx* ctx, int64_t n, const double* xyz);
int  spl_get_markers(spl_ctx* ctx, double* xyz);          /* n x 3 doubles out      */
/* Simulator.move_markers (+ Grid.move_cells under SPL_Q_MOVE_CELLS): every location
 * advances by dt * velocity of its cell; *moved = markers that changed cell.       */
int  spl_move_markers(spl_ctx* ctx, double dt, int64_t* moved);
/* Build.setup block of Simulator.advance (:66-75): Fluid = cells holding a marker,
 * air buffer of Build.max_distance = 2 rings, everything else deleted; rebuilds the
 * stencil codes / preconditioner.  sanity != 0 adds Build.check_containment and
 * Build.check_surroundings (-> SPL_E_MARKER with the reference's message).         */
int  spl_relabel(spl_ctx* ctx, int sanity);
/* current cell labels (nx*ny*nz bytes: SPL_AIR / FLUID / SOLID / ABSENT)           */
int  spl_download_media(spl_ctx* ctx, uint8_t* media);

/* keep a copy of the current u,v,w on the device / restore it (lets a bench
 * repeat the same step without a host round trip)                            */
int  spl_snapshot(spl_ctx* ctx);
int  spl_restore(spl_ctx* ctx);
/* run exactly `iters` PCG iterations in spl_project / spl_step regardless of
 * tol (0 restores the tolerance-driven loop)                                 */
int  spl_set_fixed_iters(spl_ctx* ctx, int iters);
/* iterative refinement of the pressure solve: after PCG has converged on its recursive
 * residual, r is recomputed as b - A x in fp64 from the fp64 pressure accumulator and
 * PCG is restarted on it with x kept (`rounds` times; the iterations add up in
 * spl_info.iters).  Default: 1 for fp32 contexts with SPL_PC_MG, else 0.           */
int  spl_set_refinement(spl_ctx* ctx, int rounds);
/* CUDA-event stopwatch on the context's own stream (the stream every kernel of
 * this library is launched on): start records an event, stop records a second
 * one, waits for it and returns the elapsed device time in milliseconds.      */
int  spl_timer_start(spl_ctx* ctx);
int  spl_timer_stop(spl_ctx* ctx, float* ms);
/* per-kernel timing of the PCG iteration: builds the RHS for dt, then runs
 * `iters` (<= 256) iterations with CUDA events around every phase-A, phase-B
 * and x-flush launch; returns the average duration of each in milliseconds
 * (the x flush runs once every SPL_NSBUF = 16 iterations).  The velocity field
 * is not modified (no gradient subtraction).                                   */
int  spl_profile_pcg(spl_ctx* ctx, double dt, int iters, float* ms_phase_a,
                     float* ms_phase_b, float* ms_xflush);
/* z = M^-1 r for the context's stored-z preconditioners (SPL_PC_RBIC0, SPL_PC_MG):
 * r and z are host arrays of the box (nx*ny*nz, context dtype); the labels must
 * have been uploaded.  One application of the operator PCG uses (parity tests
 * against oracle/dense_ref.py apply_rbic0 / Multigrid.vcycle); also times it:
 * *ms (nullable) receives the device time of the application.                  */
int  spl_apply_preconditioner(spl_ctx* ctx, const void* r, void* z, float* ms);
/* device pointer of a field including its z-halo (plumbing for tests):
 * which = 0 u,1 v,2 w,3 p,4 r,5 media,6 codes ; *halo_planes receives hz      */
void* spl_device_ptr(spl_ctx* ctx, int which, int* halo_planes);

#ifdef __cplusplus
}
#endif
#endif /* SPLASHY_CUDA_H */
