// b2s_internal.h -- internal types shared by the CUDA translation units of libb200sph.
// Not part of the ABI (include/b2s.h is).  Bit encodings restate the reference's contracts
// (src/source/DualSphDef.h:150-241 typecode; JDsDcellDef.h:27-43 dcell) -- they must be
// reproduced exactly for Case.bi4 / Part_XXXX.bi4 interchange and bit-exact cell tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/b2s.h"

typedef uint16_t typecode;

// ---- particle code (DualSphDef.h:150-181) ---------------------------------------------------
#define B2S_CODE_MASKSPECIAL   0xe000u
#define B2S_CODE_NORMAL        0x0000u
#define B2S_CODE_PERIODIC      0x2000u
#define B2S_CODE_OUTIGNORE     0x4000u
#define B2S_CODE_OUTMOV        0x6000u
#define B2S_CODE_OUTPOS        0x8000u
#define B2S_CODE_OUTRHO        0xA000u
#define B2S_CODE_MASKTYPE      0x1800u
#define B2S_CODE_TYPE_FIXED    0x0000u
#define B2S_CODE_TYPE_MOVING   0x0800u
#define B2S_CODE_TYPE_FLOATING 0x1000u
#define B2S_CODE_TYPE_FLUID    0x1800u
#define B2S_CODE_MASKVALUE     0x07ffu
#define B2S_CODE_SPECIAL(c)    ((c)&B2S_CODE_MASKSPECIAL)
#define B2S_CODE_TYPE(c)       ((c)&B2S_CODE_MASKTYPE)
#define B2S_CODE_TYPEVALUE(c)  ((c)&B2S_CODE_MASKVALUE)
#define B2S_CODE_SETNORMAL(c)  ((c)&(~B2S_CODE_MASKSPECIAL))
#define B2S_CODE_ISFLUID(c)    (B2S_CODE_TYPE(c)==B2S_CODE_TYPE_FLUID)
#define B2S_CODE_ISFLOATING(c) (B2S_CODE_TYPE(c)==B2S_CODE_TYPE_FLOATING)
#define B2S_CODE_ISNORMAL(c)   (B2S_CODE_SPECIAL(c)==B2S_CODE_NORMAL)

// ---- dcell (JDsDcellDef.h:32-43) ------------------------------------------------------------
#define B2S_DCEL_MAPOUT 0xffffffffu
#define B2S_DCEL_GETCODE(sx,sy,sz) ((((sx)+1)<<25)|((sy)<<20)|((sz)<<15)|(((sy)+(sz))<<10)|(((sx)+1+(sz))<<5)|((sx)+1+(sy)))
#define B2S_DCEL_CELLX(dcc,cel) ((cel)>>(((dcc)>>10)&31))
#define B2S_DCEL_CELLY(dcc,cel) (((cel)<<((dcc)>>25))>>(((dcc)>>5)&31))
#define B2S_DCEL_CELLZ(dcc,cel) (((cel)<<((dcc)&31))>>((dcc)&31))
#define B2S_DCEL_CELL(dcc,cx,cy,cz) (((cx)<<(((dcc)>>10)&31))|((cy)<<(((dcc)>>15)&31))|(cz))

#define B2S_ALMOSTZERO 1e-18f      // DualSphDef.h:114
// boundary particle modes of mDBC2 (DualSphDef.h:473-475)
#define B2S_BMODE_DBC 0
#define B2S_BMODE_MDBC2 1
#define B2S_BMODE_MDBC2OFF 2

// ---- constants passed by value to every kernel (replaces __constant__ CTE, JSphGpu_ker.cu:39) ----
struct DevCte {
  float kernelh, ovkernelh, kernelsize2, awen, bwenh;
  float cub_a2, cub_a24, cub_c1, cub_d1, cub_c2, cub_od_wdeltap;
  float cteb, gamma, ovgamma, rhopzero, ovrhopzero, cs0, eta2, massf, massb;
  float ddtkh, ddtgz, spssmag, spsblin;
  float scell, movlimit, rhopoutmin, rhopoutmax;
  float scell_blk;                 // scell*2^B2S_XBLK: x origin step of posp (see b2s_posp_x)
  float visco_f, visco_b;          // Visco and Visco*ViscoBoundFactor (JSphGpuSingle.cpp:609)
  float gravity[3];
  float ddt_c1, ddt_c2, ddt_c3;    // series coefficients of (1+x)^(1/gamma)-1
  double scell_d;                  // double(float scell): divisor of the cell hash (JSphGpu_ker.cu:387-389)
  double domposmin[3], maprealposmin[3], maprealsize[3];
  double cs0_d, kernelh_d, cfl, dtmin;
  uint32_t domcellcode, domcells[3];
  int32_t  scelldiv, tkernel, tvisco, tdensity, tboundary, tstep, use_floating, dt_allparticles, refsem, simulate2d;
  int32_t  gamma_is7;
  int32_t  mdbc2;                  // SlipMode>=2: boundary neighbours use boundmode/tangenvel (JSphCpu.cpp:851-853,990-995)
  float dp2;                       // float(Dp*Dp): scaled condition number of mDBC2
  int32_t  shiftmode; float shiftcoef, shifttfs, shiftmaxdist;   // standard shifting (JSphShifting)
  int32_t  periactive;             // periodic axes: bit0 X, bit1 Y, bit2 Z (JSph::PeriActive)
  double   perinc[3][3];           // PeriXinc, PeriYinc, PeriZinc (JSph.cpp:2186-2188)
};

// ---- device-resident state of the last cell divide (JCellDivGpu members, JCellDivGpu.h:86-138) ----
struct DivState {
  uint32_t flim[6];        // fluid  cell min(x,y,z), max(x,y,z) of valid particles (atomics)
  uint32_t blim[6];        // bound  cell limits (cached until boundary changes)
  uint32_t cellmin[3], cellmax[3];
  uint32_t ncx, ncy, ncz, nsheet, nct, nctt;
  uint32_t box_boundignore, box_fluid, box_boundout, box_fluidout, box_boundoutignore, box_fluidoutignore;
  uint32_t nptot;          // particles entering the divide
  uint32_t np, npb, npbok; // after the divide
  uint32_t npb_ignore, npb_out, npf_out, npb_outignore, npf_outignore;
  uint32_t npf_out_total;
  uint32_t error;          // !=0: boundary particle excluded etc.
  uint32_t slab_nbound_tmp; // boundary-type particles appended by the last slab unpack
  uint32_t peri_np0, peri_npb0, peri_nbper, peri_nfper;   // RunPeriodic bookkeeping (periodic.cu): Np/Npb at entry, NpbPer, NpfPer
};

// ---- device-resident step scalars (dt maxima as ordered uints for atomicMax) ----
struct StepState {
  uint32_t velmax2_bits, acemax2_bits, viscdtmax_bits, pad;
  double   dt, dt_p, dt_c, symplectic_dtpre, timestep, dt_last;
  double   velmax, acemax;
  float    viscdtmax;
  uint32_t nstep, verletstep, dtmodif;
  int32_t  error;
  unsigned long long pairs_checked, pairs_real;
};

// One set of per-particle arrays in cell-sorted order (element layouts of JSphGpu.h:125-209).
struct PartSet {
  uint32_t* idp; typecode* code; uint32_t* dcell;
  double2* posxy; double* posz; float4* velrho;
};

// One movement of a moving block (JMotionMov*, JMotionMov.h): type = B2S_MOV_*; v1,v2,v3 by type (b2s.h b2s_motion); active in
// [tstart,tend).  Movements of one block are kept together and in time order (b2s_motion_impl applies them in that order).
struct Motion {
  uint32_t idbegin, count; int32_t type, pad;
  double v1[3], v2[3], v3[3], p1[3], p2[3], tstart, tend;
};

struct b2s_ctx {
  b2s_params prm;
  DevCte cte;
  int device;
  cudaStream_t stream;
  uint32_t cap;                 // particle capacity
  uint32_t np_host, npb_host, npbok_host;   // host mirrors (upper bounds for launches)
  uint32_t npf_host;                        // upper bound of np-npb (a slab unpack raises np and npb by the same bound, so their difference is none)
  uint32_t nct_max, nctt_max;   // table capacity
  std::string err;
  uint64_t launches;
  int variant;
  // particle arrays: current + sort scratch
  PartSet cur, alt;
  // symplectic predictor copies (+scratch) / verlet m1 (+scratch)
  double2* pospre_xy; double* pospre_z; float4* velrhopre;
  double2* pospre_xy_alt; double* pospre_z_alt; float4* velrhopre_alt;
  float4* velrhom1; float4* velrhom1_alt;
  bool pre_active;
  // SPS
  float* spstau; float* spstau_alt; float* spstau_new;   // 6 floats per particle
  // mDBC
  float4* boundnor; float4* boundnor_alt;                 // xyz normal, w unused
  float4* motionace; float4* motionace_alt;               // mDBC2 with moving boundaries: acceleration of the imposed motion (MotionAce, JSphCpu_mdbc.cpp:472-490)
  float4* shiftpos;                                       // standard shifting: displacement of the current sub-step (fluid particles)
  uint8_t* boundmode; float4* tangenvel;                  // mDBC2 (SlipMode 2/3): per interaction, in the current sorted order
  // interaction inputs/outputs
  float4* posp;      // (x,y,z) relative to own cell origin, w = pressure
  uint4*  posh;      // FP16 copy of posp.xyz in cell units, two particles per 16 B: (x0,x1)(y0,y1)(z0,z1)(-) -- pre-filter of forces_pair2.cu
  float4* acear;     // ace.xyz, ar
  // neighbour list
  uint32_t* cellpart; uint32_t* sortpart; uint32_t* cellpart_alt; uint32_t* sortpart_alt;
  int2* begincell;
  bool nearbound_valid; uint8_t* nearbound;   // per cell of the cell domain: 1 = a boundary particle lies within +-scelldiv cells (interaction variant 4 skips its boundary pass otherwise)
  uint32_t* radix_counts; size_t radix_counts_len;
  uint32_t* ridp;    // index by id for moving+floating (CalcRidp)
  // device state
  DivState* div; StepState* step;
  DivState* div_h; StepState* step_h;   // pinned host mirrors
  bool bound_changed;
  uint32_t verletstep;          // JSphGpu::VerletStep (host counter)
  // floatings
  uint32_t ftcount; std::vector<b2s_floating> fts; float* ft_massp_dev; float* ft_sums_dev; float* ft_allsums_dev; float* ft_sums_h;
  void* ft_dev; // device copy of per-floating update data
  std::vector<Motion> motions; void* motion_dev; bool motion_dirty;
  // slab decomposition (slab.cu)
  void* slab_comm;              // slab_comm.cu: transport + exchange buffers (NULL until b2s_slab_comm_init*)
  bool slab_on; int slab_rank, slab_nranks; uint32_t slab_cxa, slab_cxb; uint32_t* slab_counts; uint8_t* slab_flags; size_t slab_counts_len;
  uint32_t* tile_bricks; uint32_t* tile_counters; int tile_nsm; uint32_t tile_cap;   // forces_tile.cu: list of non-empty bricks, its counters, SM count, tile capacity override (0 = default)
  // gauges (gauges.cu): device list + results of the last evaluation (device and pinned host)
  std::vector<int> gauge_types; std::vector<unsigned char> gauges_host; void* gauges_dev; float* gauge_res_dev; float* gauge_res_h; float* gauge_mass_dev; uint32_t gauge_points; float* gauge_force_dev; uint32_t gauge_force_n;
  // excluded particles (JSphGpuSingle::RunCellDivide -> AddParticlesOut -> PartsOut, JSphGpuSingle.cpp:455-470): 48-byte
  // records appended on the device by every divide, drained by b2s_out_fetch (PartOut_000.obi4); NULL = not collected
  void* out_dev; uint32_t* out_count_dev; uint32_t out_cap;
  // damping zones (damping.cu): DampOp records, host list + device copy
  std::vector<unsigned char> damp_host; void* damp_dev; bool damp_dirty;
  // scratch for byid download
  void* scratch; size_t scratch_bytes;
};

// ---- helpers -------------------------------------------------------------------------------
int  b2s_fail(b2s_ctx* ctx, int code, const char* fmt, ...);
#define B2S_CUDA(ctx, expr) do{ cudaError_t _e=(expr); if(_e!=cudaSuccess) return b2s_fail(ctx,(int)_e,"%s failed: %s (%s:%d)",#expr,cudaGetErrorString(_e),__FILE__,__LINE__); }while(0)
#define B2S_LAUNCHED(ctx) do{ (ctx)->launches++; }while(0)
static inline unsigned b2s_blocks(size_t n, unsigned bs){ return (unsigned)((n+bs-1)/bs); }

// posp frame: y,z relative to the origin of the particle's own cell; x relative to the origin of its BLOCK of 2^B2S_XBLK
// cells along x (absolute domain cells).  A whole row of neighbour cells [cx-d,cx+d] (at most 5 cells = 2^B2S_XBLK+1, so it
// touches at most two blocks) then needs at most two x offsets,
// so the interaction kernels walk a row as ONE index range instead of one range per cell.  Precision: FP32 of a length
// < 4*scell, i.e. the same order as the reference's PosCell frame (JSphGpuSimple_ker.cu:46-57).
#define B2S_XBLK 2
#ifdef __CUDACC__
__device__ __forceinline__ float b2s_offx(const DevCte& c,int cx1_abs,int cx2_abs){ return c.scell_blk*float((cx1_abs>>B2S_XBLK)-(cx2_abs>>B2S_XBLK)); }
// The position/cell part of JSphCpu::UpdatePos (JSphCpu.cpp:1471-1515) after the displacement has been applied: checks the real
// map limits, wraps the position around periodic axes, computes the cell.  Returns true when the particle is out of the map.
__device__ __forceinline__ bool b2s_pos_to_cell(const DevCte& c,double& rx,double& ry,double& rz,uint32_t& cell){
  double dx=rx-c.maprealposmin[0], dy=ry-c.maprealposmin[1], dz=rz-c.maprealposmin[2];
  bool out=(dx!=dx || dy!=dy || dz!=dz || dx<0 || dy<0 || dz<0 || dx>=c.maprealsize[0] || dy>=c.maprealsize[1] || dz>=c.maprealsize[2]);
  if(c.periactive && out){
    const bool px=(c.periactive&1)!=0, py=(c.periactive&2)!=0, pz=(c.periactive&4)!=0;
    if(px){
      if(dx<0)                { dx-=c.perinc[0][0]; dy-=c.perinc[0][1]; dz-=c.perinc[0][2]; }
      if(dx>=c.maprealsize[0]){ dx+=c.perinc[0][0]; dy+=c.perinc[0][1]; dz+=c.perinc[0][2]; }
    }
    if(py){
      if(dy<0)                { dx-=c.perinc[1][0]; dy-=c.perinc[1][1]; dz-=c.perinc[1][2]; }
      if(dy>=c.maprealsize[1]){ dx+=c.perinc[1][0]; dy+=c.perinc[1][1]; dz+=c.perinc[1][2]; }
    }
    if(pz){
      if(dz<0)                { dx-=c.perinc[2][0]; dy-=c.perinc[2][1]; dz-=c.perinc[2][2]; }
      if(dz>=c.maprealsize[2]){ dx+=c.perinc[2][0]; dy+=c.perinc[2][1]; dz+=c.perinc[2][2]; }
    }
    const bool outx=!px && (dx<0 || dx>=c.maprealsize[0]);
    const bool outy=!py && (dy<0 || dy>=c.maprealsize[1]);
    const bool outz=!pz && (dz<0 || dz>=c.maprealsize[2]);
    out=(outx||outy||outz);
    rx=dx+c.maprealposmin[0]; ry=dy+c.maprealposmin[1]; rz=dz+c.maprealposmin[2];
  }
  if(out){ cell=B2S_DCEL_MAPOUT; return true; }
  if(c.periactive){ dx=rx-c.domposmin[0]; dy=ry-c.domposmin[1]; dz=rz-c.domposmin[2]; }
  cell=B2S_DCEL_CELL(c.domcellcode,uint32_t(dx/c.scell_d),uint32_t(dy/c.scell_d),uint32_t(dz/c.scell_d));
  return false;
}
__device__ __forceinline__ float b2s_press(float rho, const DevCte& c){
  // Tait EOS as the CPU path evaluates it (FunSphEos.h:37-39): cteb*(powf(rho/rho0,gamma)-1).
  // gamma==7: exact product chain in double, rounded once (== correctly rounded powf to <1ulp).
  const float r=rho*c.ovrhopzero;
  float pw;
  if(c.gamma_is7){ const double d=(double)r; const double d2=d*d; const double d4=d2*d2; pw=(float)(d4*d2*d); }
  else pw=(float)pow((double)r,(double)c.gamma);
  return c.cteb*(pw-1.0f);
}
#endif

#ifdef __CUDACC__
// fsph::GetKernel_WabFac (FunSphKernel.h:71-85 cubic, :235-245 Wendland)
__device__ __forceinline__ float wab_fac(int tkernel,float rr2,const DevCte& c,float& fac){
  const float rad=sqrtf(rr2);
  const float qq=rad/c.kernelh;
  if(tkernel==B2S_KERNEL_WENDLAND){
    const float wqq1=1.f-0.5f*qq;
    const float wqq2=wqq1*wqq1;
    fac=c.bwenh*wqq2*wqq1;
    const float wqq=qq+qq+1.f;
    return c.awen*wqq*wqq2*wqq2;
  }
  if(rad>c.kernelh){
    const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1;
    fac=(c.cub_c2*wqq2/rad);
    return c.cub_a24*(wqq2*wqq1);
  }
  const float wqq2=qq*qq;
  fac=((c.cub_c1*qq+c.cub_d1*wqq2)/rad);
  return c.cub_a2*(1.0f+(0.75f*qq-1.5f)*wqq2);
}
#endif

// implemented in the .cu files
int b2s_celldiv_impl(b2s_ctx* ctx);
int b2s_forces_impl(b2s_ctx* ctx, int interstep);
int b2s_integrate_impl(b2s_ctx* ctx, int mode, double dt, bool dt_from_device);   // mode: 0 verlet,1 sympre,2 symcorr
int b2s_mdbc_impl(b2s_ctx* ctx);
int b2s_shifting_impl(b2s_ctx* ctx);   // shifting.cu: RunShifting with StepState.dt, called by b2s_integrate_impl (corrector / Verlet)
int b2s_periodic_impl(b2s_ctx* ctx);   // periodic.cu: RunPeriodic (duplicates behind the current particles; no-op without periodic axes)
int b2s_damping_impl(b2s_ctx* ctx);    // damping.cu: RunDamping with StepState.dt
int b2s_floating_impl(b2s_ctx* ctx, double dt, bool predictor, bool dt_from_device);
int b2s_motion_impl(b2s_ctx* ctx, double timestep, double dt);
int b2s_sync_info(b2s_ctx* ctx);
int b2s_gauges_run(b2s_ctx* ctx, int phase);   // gauges.cu: evaluates every gauge on the current particle state (0: one context; 1, 2: around the all-reduce of a slab run)
