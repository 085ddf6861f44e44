/* b2s.h -- C-ABI of libb200sph: the B200-native WCSPH time-step hot path.
 *
 * Drop-in boundary: the reference calls its GPU kernels through free
 * functions in the namespaces cusph:: / cusphs:: / cudiv:: taking POD structs and raw device
 * pointers (all file:line below are relative to /root/reference/src/source).  This header
 * declares the coarser, fusable entry points that replace them; every function cites the
 * reference interface it stands in for.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = b2s error (B2S_ERR_*), >0 = cudaError_t.
 *   - b2s_last_error(ctx) returns a static/ctx-owned message; no exception crosses the ABI.
 *   - one ctx per GPU; a ctx is not thread-safe, different ctxs are independent.
 *   - the ctx owns all device memory (the reference's JArraysGpu/JCellDivGpu pools,
 *     JArraysGpu.h, JCellDivGpu.h:86-138); callers see host copies through the getters.
 *   - there is NO CPU fallback: without a CUDA device b2s_create fails with
 *     B2S_ERR_NODEVICE and nothing else can be called.
 */
#ifndef B2S_H_
#define B2S_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 100

/* error codes */
#define B2S_OK              0
#define B2S_ERR_ARG        -1
#define B2S_ERR_NODEVICE   -2
#define B2S_ERR_NOMEM      -3
#define B2S_ERR_STATE      -4
#define B2S_ERR_UNSUPPORTED -5
#define B2S_ERR_DTINVALID  -6   /* dt is NaN/Inf: JSphGpu.cpp:1122-1129 */
#define B2S_ERR_BOUNDOUT   -7   /* boundary particle left the domain: JSphGpuSingle.cpp:505 */
#define B2S_ERR_IO         -8
#define B2S_ERR_NAME       -9

/* enums: same numeric values as the reference (DualSphDef.h:435-566, JViscosityDef.h) */
enum { B2S_KERNEL_CUBIC = 1, B2S_KERNEL_WENDLAND = 2 };                    /* TpKernel   */
enum { B2S_STEP_VERLET = 1, B2S_STEP_SYMPLECTIC = 2 };                     /* TpStep     */
enum { B2S_VISCO_ARTIFICIAL = 1, B2S_VISCO_LAMINARSPS = 2, B2S_VISCO_LAMINAR = 3 }; /* TpVisco */
enum { B2S_DDT_NONE = 0, B2S_DDT_MOLTENI = 1, B2S_DDT_FOURTAKAS = 2, B2S_DDT_FOURTAKASFULL = 3 }; /* TpDensity */
enum { B2S_BC_DBC = 1, B2S_BC_MDBC = 2 };                                  /* TpBoundary */
enum { B2S_SLIP_VEL0 = 1, B2S_SLIP_NOSLIP = 2, B2S_SLIP_FREESLIP = 3 };    /* TpSlipMode */
enum { B2S_SHIFT_NONE = 0, B2S_SHIFT_NOBOUND = 1, B2S_SHIFT_NOFIXED = 2, B2S_SHIFT_FULL = 3 };  /* TpShifting (DualSphDef.h) */
enum { B2S_INTER_VERLET = 1, B2S_INTER_SYMPREDICTOR = 2, B2S_INTER_SYMCORRECTOR = 3 }; /* TpInterStep */
/* Which reference build a semantic detail follows where CPU and GPU differ
 * (DDT2 near DBC boundaries, JSphCpu.cpp:1318-1340 vs JSphGpu_ker.cu:939-946). */
enum { B2S_REFSEM_CPU = 0, B2S_REFSEM_GPU = 1 };

/* Constants + configuration of one simulation.  Replaces StCteInteraction (JSphGpu_cte.h:27-63)
 * + the members of JSph that ConfigConstants1/2, ConfigCellDivision, SelecDomain fill
 * (JSph.cpp:1481-1546,1877-1946).  Filled by b2s_config_finalize() from the primary values. */
typedef struct b2s_params {
  /* ---- primary values (case XML: constants + parameters) ---- */
  int32_t  simulate2d;
  int32_t  tkernel, tstep, tvisco, tdensity, tboundary, slipmode;
  int32_t  verletsteps;
  int32_t  cellmode_half;      /* 0: cells of 2h (CELLMODE_Full), 1: cells of h (CELLMODE_Half) */
  int32_t  dt_allparticles;
  int32_t  mdbc_corrector;
  int32_t  refsem;             /* B2S_REFSEM_* */
  int32_t  use_floating;       /* casenfloat>0 */
  int32_t  shiftmode;          /* TpShifting: 0 none, 1 NoBound, 2 NoFixed, 3 Full (JSph.cpp:744-758) */
  float    shiftcoef, shifttfs; /* ShiftCoef (default -2), ShiftTFS (0: no free-surface threshold) */
  float    kernelh, cteb, gamma, rhopzero, massfluid, massbound;
  float    visco, viscoboundfactor, ddtvalue;
  float    rhopoutmin, rhopoutmax;
  float    gravity[3];
  double   dp, cflnumber, dtini, dtmin, coefdtmin, ftpause;
  double   maprealposmin[3], maprealposmax[3];   /* after ResizeMapLimits (JSph.cpp:1445-1476) */
  uint32_t casenp, casenfixed, casenmoving, casenfloat, casenfluid;
  /* periodic boundaries (JSph.cpp:815-827): bit0 X, bit1 Y, bit2 Z; periNinc = displacement applied to a particle that
   * leaves through the upper face of axis N.  Primary input: the off-axis components (XPeriodicIncY ...); the component
   * along the axis itself is set by b2s_config_finalize to -MapRealSize (JSph.cpp:2186-2188).  maprealposmin/max of a
   * periodic axis = case limits -+ dp/2 (JPartsLoad4::CalculeLimitsPos, JPartsLoad4.cpp:389-399). */
  int32_t  periactive;
  double   perixinc[3], periyinc[3], perizinc[3];
  /* ---- derived values (b2s_config_finalize) ---- */
  float    kernelsize, kernelsize2, awen, bwenh;            /* FunSphKernel.h:198-209 */
  float    cubic_a1, cubic_a2, cubic_aa, cubic_a24, cubic_c1, cubic_d1, cubic_c2, cubic_od_wdeltap; /* :45-85 */
  float    eta2, ddtkh, ddtgz, spssmag, spsblin, scell, movlimit, ovrhopzero;
  double   cs0;
  double   map_posmin[3], map_posmax[3], map_size[3];       /* == maprealpos; widened by KernelSize on periodic axes (JSph.cpp:2191-2195) */
  double   maprealsize[3];
  uint32_t map_cells[3];                                     /* ceil(map_size/scell) */
  uint32_t domcells[3];                                      /* single domain: == map_cells */
  uint32_t domcellcode;                                      /* DCEL_GetCode (JDsDcellDef.h:32) */
  int32_t  scelldiv;                                         /* 1 full, 2 half */
  double   domposmin[3], domrealposmin[3], domrealposmax[3];
  float    poscellsize;                                      /* JSph.cpp:1949-1990 */
  uint32_t casenbound, casenpb;
} b2s_params;

/* Floating body data: StFloatingData (DualSphDef.h) subset used by FtComputeAceVel (JSph.cpp:2565-2704) */
typedef struct b2s_floating {
  uint32_t begin, count;        /* id range of its particles */
  float    mass, massp, radius;
  int32_t  constraints;         /* bit mask: 1 movx,2 movy,4 movz,8 rotx,16 roty,32 rotz blocked */
  double   center[3];
  float    angles[3];
  float    fvel[3], fomega[3];
  float    facelin[3], faceang[3];
  float    inertiaini[9];       /* row-major 3x3 */
  /* values of the last corrector / Verlet update, for PartFloatInfo.ibi4 (JSph.cpp:2613-2620): force and torque of the
   * fluid on the body (N, N*m) and the accelerations before the motion constraints (gravity included) */
  float    fluforcelin[3], fluforceang[3], preacelin[3], preaceang[3];
} b2s_floating;

/* Run-time counters and maxima, one coherent snapshot (replaces the blocking scalar
 * cudaMemcpy calls of JSphGpu_ker.cu:117,178 and JCellDivGpu.cpp:297-309). */
typedef struct b2s_info {
  uint32_t np, npb, npbok, npf_out_total;
  uint32_t ncx, ncy, ncz, nct;
  uint32_t cellmin[3], cellmax[3];
  uint32_t npb_ignore, npf_out_last, npb_out_last, nstep;
  double   timestep, dt_last, dt_p, dt_c, symplectic_dtpre;
  double   velmax, acemax;
  float    viscdtmax, pad0;
  uint32_t verletstep, dtmodif;
  uint64_t pairs_checked, pairs_real;   /* filled only by b2s_count_pairs */
} b2s_info;

typedef struct b2s_ctx b2s_ctx;

/* ---- configuration (host math; no device needed) ---------------------------------------- */
/* Fills the derived block of *p from its primary block.  Mirrors JSph::ConfigConstants1/2,
 * ConfigCellDivision, SelecDomain, ConfigPosCellGpu, JDsDcell::CalcCellCode
 * (JSph.cpp:1481-1546,1877-1990; JDsDcell.cpp:41-67). */
int b2s_config_finalize(b2s_params* p);
const char* b2s_error_string(int code);

/* ---- context ------------------------------------------------------------------------------ */
/* device<0: current device.  capacity_np: particle slots to allocate (>= np + halo slack). */
int b2s_create(const b2s_params* p, int device, uint32_t capacity_np, b2s_ctx** out);
int b2s_destroy(b2s_ctx* ctx);
const char* b2s_last_error(const b2s_ctx* ctx);
int b2s_get_params(const b2s_ctx* ctx, b2s_params* out);

/* ---- particle data up / down (JSphGpu::ParticlesDataUp/Down, JSphGpu.cpp:606-686) -------- */
/* Host arrays in the reference's element layouts: idp u32, code u16 (typecode), pos double[3],
 * velrho float[4]; boundnor float[3] optional (NULL unless mDBC).  Particles must be ordered
 * [bound(fixed,moving) | floating | fluid] as JPartsLoad4 guarantees (JPartsLoad4.cpp:131-140).
 * Computes dcell (JSph::LoadDcellParticles, JSph.cpp:1794-1820) and runs the first cell divide
 * like JSphGpuSingle::ConfigDomain (JSphGpuSingle.cpp:151-240). */
int b2s_upload_particles(b2s_ctx* ctx, uint32_t np, const uint32_t* idp, const uint16_t* code,
                         const double* pos, const float* velrho, const float* boundnor);
/* Copies the current state (cell-sorted order) to host arrays (any may be NULL). Returns np in *np_out. */
int b2s_download_particles(b2s_ctx* ctx, uint32_t* np_out, uint32_t* idp, uint16_t* code,
                           double* pos, float* velrho);
/* Same but ordered by idp and only non-excluded particles: what SaveData writes (JSphGpuSingle.cpp:1053). */
int b2s_download_particles_byid(b2s_ctx* ctx, uint32_t* np_out, uint32_t* idp, double* pos,
                                float* vel, float* rhop);

/* Named read-back of any ctx-owned device array (test entry point; replaces the debug dumps of
 * JDebugSphGpu).  Names: idp code dcell pos velrho cellpart sortpart begincell(int2 table)
 * ace ar press poscell velrhopre pospre velrhom1 spstau boundnor.  *count_out = elements copied;
 * fails with B2S_ERR_ARG if capacity_bytes is too small. */
int b2s_get_array(b2s_ctx* ctx, const char* name, void* dst, size_t capacity_bytes, size_t* count_out);
int b2s_get_info(b2s_ctx* ctx, b2s_info* out);
/* Positions of n particles of the moving / floating blocks by id (CaseNfixed <= id < CaseNbound), through the device-side
 * id -> index table of the last divide (RidpMot, JSphGpu.cpp:1173-1180): what JDsPartMotionSave::AddDataExtraGpu
 * (JDsPartMotionSave.cpp:301-313) reads for the extra-frequency motion file.  pos_out[3n]; a particle that left the
 * simulation comes back as NaN. */
int b2s_get_pos_byid(b2s_ctx* ctx, uint32_t n, const uint32_t* ids, double* pos_out);
/* Restart support (JSph.cpp:2214-2225, -partbegin): sets the simulated time and, if >0, SymplecticDtPre (the dt the next
 * Symplectic step starts with).  Call after b2s_upload_particles. */
int b2s_set_time(b2s_ctx* ctx, double timestep, double symplectic_dtpre);

/* ---- neighbour list: JSphGpuSingle::RunCellDivide (JSphGpuSingle.cpp:372-498) ------------ */
/* = cudiv::LimitsCell + PreSortFull/Fluid + Sort + CalcBeginEndCell + SortDataParticles +
 *   cusphs::UpdatePosCell (+CalcRidp)  (JCellDivGpu_ker.cu:439,120,533,577; JCellDivGpuSingle_ker.cu:41;
 *   JSphGpuSimple_ker.cu:41). */
int b2s_celldiv(b2s_ctx* ctx);

/* ---- interaction: PreInteraction_Forces + cusph::Interaction_Forces + reductions ---------- */
/* (JSphGpu.cpp:846-913, JSphGpu_ker.cu:1043-1179, JSphGpuSingle.cpp:590-727).  interstep =
 * B2S_INTER_*.  Runs mDBC correction first when tboundary==MDBC (JSphGpuSingle.cpp:686-712),
 * then fluid and boundary force kernels, SPS tau, and the fused max reductions; computes
 * dt = DtVariable() on the device (JSphGpu.cpp:1103-1147). */
int b2s_interaction(b2s_ctx* ctx, int interstep);

/* ---- integrators (cusphs::ComputeStepVerlet/SymplecticPre/Cor + cusph::ComputeStepPos[2]) - */
/* (JSphGpuSimple_ker.cu:132-553, JSphGpu_ker.cu:1476-1570; callers JSphGpu.cpp:939-1097). */
int b2s_compute_verlet(b2s_ctx* ctx, double dt);
int b2s_compute_symplectic_pre(b2s_ctx* ctx, double dt);
int b2s_compute_symplectic_corr(b2s_ctx* ctx, double dt);

/* ---- whole step: JSphGpuSingle::ComputeStep + the Run-loop tail (JSphGpuSingle.cpp:761-819,998-1010)
 * = ComputeStep_Sym/Ver, RunMotion for registered linear motions, RunCellDivide, TimeStep+=dt.
 * One host synchronisation per call (to return dt).  nsteps>1 runs several steps back to back. */
int b2s_compute_step(b2s_ctx* ctx, double* dt_out);
int b2s_run_steps(b2s_ctx* ctx, uint32_t nsteps, double* time_out);

/* ---- moving boundaries: cusph::MoveLinBound (JSphGpu_ker.cu:1671-1700) via RunMotion (JSphGpu.cpp:1173-1225)
 * Registers a sinusoidal rectilinear motion (JMotionMovRectSinu) for the id block [idbegin,idbegin+count):
 * pos(t)=ampl*sin(2*pi*freq*t+phase) per axis; applied every step with the step's dt. */
int b2s_add_motion_rectsinu(b2s_ctx* ctx, uint32_t idbegin, uint32_t count, const double freq[3],
                            const double ampl[3], const double phase[3], double tstart, double tend);
/* General form: one movement of JMotion (JMotionMov.h:106-276, evaluated like JMotionObj::ProcesTime, JMotionObj.cpp:470-600)
 * for the moving block [idbegin,idbegin+count), active in [tstart,tend).  Movements of one block are applied in the order they
 * were added; linear results go through cusph::MoveLinBound, results with a rotation through cusph::MoveMatBound
 * (JSphGpu_ker.cu:1671-1759).  Angles in degrees (as the reference stores them), phases in radians.
 *   type            v1                 v2               v3
 *   RECTSINU        freq (Hz)          ampl (m)         phase (rad)
 *   RECT            vel (m/s)
 *   RECTACE         velini (m/s)       ace (m/s^2)
 *   ROT             v1[0]=velang (deg/s)                                  axisp1, axisp2
 *   ROTACE          v1[0]=velini       v2[0]=aceang (deg/s^2)             axisp1, axisp2
 *   ROTSINU         v1[0]=freq (Hz)    v2[0]=ampl (deg) v3[0]=phase (rad) axisp1, axisp2
 *   WAIT            (nothing moves) */
enum { B2S_MOV_RECTSINU = 0, B2S_MOV_RECT = 1, B2S_MOV_RECTACE = 2, B2S_MOV_ROT = 3, B2S_MOV_ROTACE = 4, B2S_MOV_ROTSINU = 5, B2S_MOV_WAIT = 6 };
typedef struct b2s_motion {
  uint32_t idbegin, count;
  int32_t  type, pad;
  double   tstart, tend;
  double   v1[3], v2[3], v3[3];
  double   axisp1[3], axisp2[3];
} b2s_motion;
int b2s_add_motion(b2s_ctx* ctx, const b2s_motion* m);
/* Direct form: displace a block now (MOTT_Linear data: JDsMotionDef.h). */
int b2s_move_linear(b2s_ctx* ctx, uint32_t idbegin, uint32_t count, const double mvpos[3], const float mvvel[3]);

/* ---- excluded particles: JSphGpuSingle::RunCellDivide -> ParticlesDataDown(...) -> AddParticlesOut (JSphGpuSingle.cpp:455-470,
 * JSph.cpp:2981-3040): after b2s_out_enable every cell divide appends the fluid particles it excludes (id, position, velocity,
 * density, motive 1 position / 2 density / 3 displacement) to a device buffer; b2s_out_fetch drains it (PartOut_000.obi4). */
int b2s_out_enable(b2s_ctx* ctx, uint32_t capacity);
int b2s_out_fetch(b2s_ctx* ctx, uint32_t maxcount, uint32_t* count_out, uint32_t* idp, double* pos, float* vel, float* rhop, uint8_t* motive);

/* ---- damping zones: JDsDamping (JDsDamping.cpp:96-131,185-233 plane; :345-391,446-492 box), JSphCpu::RunDamping
 * (JSphCpu.cpp:2312-2316).  Registered zones are applied to the fluid particles once per step, after the integrator and the
 * floating-body update (JSphCpuSingle.cpp:885,928), in registration order.
 * plane: <dampingzone> limitmin -> limitmax gives the direction and the ramp length, overlimit extends it; domain = NULL or
 * 10 doubles {zmin, zmax, x1,y1, x2,y2, x3,y3, x4,y4} (<domain>).  box: <dampingbox> with the four corner points and the
 * active directions (B2S_DAMPDIR_* mask). */
enum { B2S_DAMPDIR_TOP = 1, B2S_DAMPDIR_BOTTOM = 2, B2S_DAMPDIR_LEFT = 4, B2S_DAMPDIR_RIGHT = 8, B2S_DAMPDIR_FRONT = 16,
       B2S_DAMPDIR_BACK = 32, B2S_DAMPDIR_ALL = 63 };   /* JDsDampingOp_Box::TpDirections (JDsDamping.h:171-181) */
int b2s_add_damping_plane(b2s_ctx* ctx, const double limitmin[3], const double limitmax[3], float overlimit, float redumax,
                          const float factorxyz[3], const double* domain);
int b2s_add_damping_box(b2s_ctx* ctx, const double limitmin_ini[3], const double limitmin_end[3], const double limitmax_ini[3],
                        const double limitmax_end[3], float overlimit, float redumax, const float factorxyz[3], uint32_t directions);

/* ---- floating bodies: FtPartsSumAce / FtComputeAceVel / FtPartsUpdate ---------------------- */
/* (JSphGpu_ker.cu:1938-2101, JSph.cpp:2565-2704, JSphGpuSingle.cpp:827-907). */
int b2s_set_floatings(b2s_ctx* ctx, uint32_t ftcount, const b2s_floating* fts);
int b2s_get_floatings(b2s_ctx* ctx, uint32_t ftcount, b2s_floating* fts);

/* ---- gauges: JGaugeVelocity / JGaugeSwl (JDsGaugeItem.cpp:427-478,728-790), evaluated on the device after every step at the
 * point of the step where the reference calls RunGaugeSystem (JSphCpuSingle.cpp:1192).  The timing rules (computedt/outputdt)
 * and the Gauges*.csv files are the host driver's (b2s_host.h); b2s_gauge_result returns the last evaluation:
 * velocity (vx,vy,vz) or the surface point (x,y,z).  b2s_gauges_compute evaluates them now (initial measurement at t=0). */
int b2s_gauge_add_velocity(b2s_ctx* ctx, const double point[3], int* id);
int b2s_gauge_add_swl(b2s_ctx* ctx, const double point0[3], const double point2[3], double pointdp, float masslimit, int* id);
/* Force gauge (JGaugeForce, JDsGaugeItem.cpp:1847-1908; JGaugeSystem::AddGaugeForce, JDsGaugeSystem.cpp:472-495): pressure force
 * of the fluid on the fixed or moving boundary block with ids [idbegin,idbegin+count) whose particles carry `typevalue` in the
 * type+value bits of their code; result = (fx,fy,fz) in N. */
int b2s_gauge_add_force(b2s_ctx* ctx, uint32_t idbegin, uint32_t count, uint32_t typevalue, int* id);
/* MaxZ gauge (JGaugeMaxZ, JDsGaugeItem.cpp:1023-1091): highest fluid particle within the horizontal distance distlimit of the
 * vertical line point0 .. point0+height; result[0] = zmax (never below the domain's z minimum). */
int b2s_gauge_add_maxz(b2s_ctx* ctx, const double point0[3], double height, float distlimit, int* id);
int b2s_gauges_compute(b2s_ctx* ctx);
int b2s_gauge_result(b2s_ctx* ctx, int id, float out[3]);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* Counts candidate / real (r<2h) pairs of the current neighbour list (what JDsPips reports,
 * JDsPips.cpp:116-142) for the flop model. */
int b2s_count_pairs(b2s_ctx* ctx, uint64_t* checked, uint64_t* real);
/* Times `reps` launches of the fluid interaction kernel alone on the ctx stream with CUDA events;
 * returns mean milliseconds per launch (used by bench.py's roofline leg). */
int b2s_time_interaction(b2s_ctx* ctx, int reps, float* ms_per_launch);
/* Number of kernel launches issued by this ctx since creation / since last reset. */
int b2s_launch_count(b2s_ctx* ctx, uint64_t* launches, int reset);
/* CUDA stream the ctx launches on (cudaStream_t cast to void*), for external event timing. */
int b2s_get_stream(b2s_ctx* ctx, void** stream_out);
/* Selects the fluid interaction kernel variant:
 * 5 = (default) FP16 pre-filter into per-lane hit masks, two hits per lane and round in packed FP32, mirror rows
 *     (csrc/forces_pair2.cu); instantiated for every axis of the reference's dispatch (JSphGpu_ker.cu:1080-1179: Wendland / cubic,
 *     artificial / laminar / laminar+SPS, every DDT mode, floating bodies, mDBC2); equal to 0 within rounding: the pair terms are
 *     re-associated (csrc/forces_pair2.cuh);
 * 0 = gather: one thread per particle, the reference's summation order (the parity anchor of the golden fixtures);
 * 6 = bricks of cells staged in shared memory in one FP32 frame per tile, per-lane hit lists over many rows, the same packed
 *     pair body (csrc/forces_tile.cu; Wendland + artificial viscosity without floating bodies, otherwise runs 5; measured slower
 *     than 5, kept as the record of the shared-memory design: profiles/README.md).
 * Other values: B2S_ERR_ARG. */
int b2s_set_interaction_variant(b2s_ctx* ctx, int variant);
/* Test hook for variant 6: particle capacity of a shared-memory tile (0 = built-in 4096).  A lower value makes small cases take
 * the column split and the global-memory fallback of csrc/forces_tile.cu.  Results do not depend on it beyond rounding. */
int b2s_debug_set_tile_cap(b2s_ctx* ctx, uint32_t cap);

/* ---- slab decomposition over several GPUs (no reference counterpart) ----------- */
/* One ctx per GPU owns the particles whose cell-x lies in [cx_begin,cx_end) of the GLOBAL cell grid.  Before every
 * interaction the owner ranks exchange (1) migrants = owned particles whose cell left the slab (full state) and
 * (2) halos = owned particles within scelldiv cell layers of a cut (appended on the neighbour as CODE_PERIODIC
 * ghosts, like the reference's periodic duplicates, JSphGpu_ker.cu:2267-2399).  The transport is the caller's
 * (ncclSend/ncclRecv through torch.distributed, see dualsphysics_b200/slab.py); these entry points only pack into /
 * unpack from contiguous DEVICE buffers, in deterministic particle order.
 * kind: 0 = halo, 1 = migrants.  rank 0 has no left neighbour, rank nranks-1 no right neighbour. */
int b2s_slab_config(b2s_ctx* ctx, int rank, int nranks, uint32_t cx_begin, uint32_t cx_end);
int b2s_slab_record_bytes(b2s_ctx* ctx, int kind, uint32_t* bytes);
int b2s_slab_pack(b2s_ctx* ctx, int kind, void* dev_left, void* dev_right, uint32_t cap_records,
                  uint32_t* nleft, uint32_t* nright);
int b2s_slab_unpack(b2s_ctx* ctx, int kind, const void* dev_buf, uint32_t count);
/* Merged exchange of one interaction: migrants and halos in ONE message per neighbour.  counts = {migrants left, migrants right,
 * halo left, halo right}; each buffer holds the migrant records followed by the halo records.  Migrated particles stay on the
 * sender as halo ghosts of their new owner.  Needs slabs wider than the halo. */
int b2s_slab_pack_both(b2s_ctx* ctx, void* dev_left, void* dev_right, uint64_t cap_bytes, uint32_t counts[4]);
int b2s_slab_unpack_both(b2s_ctx* ctx, const void* dev_buf, uint32_t nmigrants, uint32_t nhalo);
/* Device address of the three dt maxima {velmax^2, acemax^2, viscdtmax} stored as order-preserving uint32 bit
 * patterns (non-negative floats): all-reduce them with MAX (ncclMax on int32) between b2s_slab_forces and
 * b2s_slab_apply_dt so that every rank computes the same DtVariable (JSphGpu.cpp:1103-1147). */
int b2s_slab_dt_buffer(b2s_ctx* ctx, void** dev_u32x3);
/* The step of JSphGpuSingle::ComputeStep_Sym/Ver cut at the points where ranks must talk:
 *   b2s_slab_forces     = MdbcBoundCorrection + Interaction_Forces without the dt computation
 *   b2s_slab_apply_dt   = DtVariable from the (all-reduced) maxima; interstep as in b2s_interaction
 *   b2s_slab_integrate  = mode 0 ComputeVerlet, 1 ComputeSymplecticPre, 2 ComputeSymplecticCorr with the device dt
 *   b2s_slab_step_end   = SymplecticDtPre=min(dt_p,dt_c), TimeStep+=dt, Nstep++ ; returns dt (one host sync) */
/* Re-cutting (dynamic load balance): b2s_slab_histogram counts the owned particles per x cell plane of the GLOBAL grid into
 * dev_u32_ncx[domcells[0]] (all-reduce it with SUM, derive new cuts), b2s_slab_set_range installs the new [cx_begin,cx_end);
 * the particles now outside leave through the normal migrant exchange (repeat it until no rank has migrants left). */
int b2s_slab_histogram(b2s_ctx* ctx, void* dev_u32_ncx);
int b2s_slab_set_range(b2s_ctx* ctx, uint32_t cx_begin, uint32_t cx_end);
int b2s_slab_get_range(b2s_ctx* ctx, uint32_t* range2);   /* current [cx_begin,cx_end) of this rank (changes with a re-cut) */
/* Marks the halo ghosts of the last exchange for removal by the next divide (needed when halos are rebuilt between steps). */
int b2s_slab_drop_ghosts(b2s_ctx* ctx);
int b2s_slab_forces(b2s_ctx* ctx, int interstep);
int b2s_slab_apply_dt(b2s_ctx* ctx, int interstep);
int b2s_slab_integrate(b2s_ctx* ctx, int mode);
int b2s_slab_step_end(b2s_ctx* ctx, double* dt_out);
/* Floating bodies across slabs (RunFloating, JSphGpuSingle.cpp:827-907, cut where the ranks must talk): after
 * b2s_slab_integrate call b2s_slab_ft_sums (force/torque sums of the floating particles THIS rank owns; returns the device
 * buffer of 6*ftcount floats), all-reduce that buffer with SUM, then b2s_slab_ft_apply (rigid-body algebra, identical on
 * every rank, and update of the owned floating particles).  predictor: 1 after SymplecticPre, 0 otherwise. */
int b2s_slab_ft_sums(b2s_ctx* ctx, void** dev_f32, uint32_t* count);
int b2s_slab_ft_apply(b2s_ctx* ctx, int predictor);
/* ---- the slab step inside the library (csrc/slab_comm.cu): NCCL transport, no host synchronisation inside a step ----------
 * One process per GPU.  Rank 0 makes the 128-byte NCCL id (b2s_slab_comm_id) and hands it to every rank by whatever means the
 * launcher has (torch.distributed broadcast in bench.py / slab.py, a file in the CLI); every rank then calls b2s_slab_comm_init
 * after b2s_slab_config.  max_plane_particles (b2s_sim_slab_info) sizes the exchange buffers.  NCCL is dlopen'ed here, so
 * single-GPU runs do not need it.  Collective calls (every rank, same order):
 *   b2s_slab_exchange   migrants + halos to both neighbours in one fixed-size message per side whose header carries the counts
 *                       (they never visit the host), then the cell divide; call once after b2s_slab_comm_init
 *   b2s_slab_step       one whole ComputeStep (JSphGpuSingle.cpp:761-819 + run-loop tail :1003-1008): interactions with the dt
 *                       maxima all-reduced (MAX), integrators, floating-body sums all-reduced (SUM), the exchanges, an optional
 *                       re-cut; ONE host synchronisation at its end
 *   b2s_slab_recut      dynamic load balance: all-reduced work histogram along x -> new cut planes -> migrant rounds -> halos
 * b2s_slab_comm_set_recut: b2s_slab_step re-cuts every that many steps (0 = never).
 * b2s_slab_comm_stats: out[0] ms inside pack + transfer + unpack (events on the context's stream: includes waiting for a slower
 *   neighbour), [1] exchanges, [2] bytes sent, [3] re-cuts since the last reset, [4] migrants sent since b2s_slab_comm_init.
 * b2s_slab_comm_init_local: test transport -- all ranks in this process on one device (device copies, host reductions); one call
 *   of a collective entry point on any context of the group then drives every rank. */
int b2s_slab_comm_id(uint8_t* id128);
int b2s_slab_comm_init(b2s_ctx* ctx, const uint8_t* id128, uint32_t max_plane_particles);
int b2s_slab_comm_init_local(b2s_ctx** ctxs, int nranks, uint32_t max_plane_particles);
int b2s_slab_comm_free(b2s_ctx* ctx);
int b2s_slab_comm_set_recut(b2s_ctx* ctx, uint32_t every_steps);
int b2s_slab_comm_stats(b2s_ctx* ctx, double* out5, int reset);
/* Gauges of a slab-decomposed run, evaluated collectively (every rank calls it; all ranks end with the same b2s_gauge_result): the
 * rank whose slab holds a measurement point evaluates it, force gauges sum over the owners of the boundary particles, maxz gauges
 * take the maximum over the ranks.  b2s_slab_step does the same at the end of every step; this call is for the measurement before
 * the first step (JSphCpuSingle.cpp:1110-1111).  The same gauges have to be registered on every rank. */
int b2s_slab_gauges_compute(b2s_ctx* ctx);
/* B2S_SLAB_PROFILE=1 in the environment: ms per step spent in every phase of b2s_slab_step (CUDA events at the phase boundaries,
 * collected at the step's synchronisation) as one line of text; empty otherwise. */
int b2s_slab_comm_profile(b2s_ctx* ctx, char* out, uint32_t outlen);
int b2s_slab_exchange(b2s_ctx* ctx);
int b2s_slab_step(b2s_ctx* ctx, double* dt_out);
int b2s_slab_run_steps(b2s_ctx* ctx, uint32_t nsteps, double* time_out);
int b2s_slab_recut(b2s_ctx* ctx, uint32_t* range_out, uint32_t n2);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H_ */
