// ctx.cu -- context, device memory, data up/down, and the step driver of libb200sph (C-ABI in include/b2s.h).
//
// Step orchestration mirrors JSphGpuSingle::ComputeStep_Sym / ComputeStep_Ver and the tail of the Run loop
// (src/source/JSphGpuSingle.cpp:761-819, 998-1010); the CPU twin is JSphCpuSingle.cpp:871-932,1188-1196.
// Differences by design (B200-first): every reduction result, dt and the cell-domain state stay on the device
// (DivState / StepState), kernels are enqueued back to back on one stream and the host synchronises ONCE per
// step to read {dt, counters, error flags} from pinned memory (the reference blocks >=6 times per step:
// JSphGpu_ker.cu:117,178; JCellDivGpu.cpp:297-309; JSphGpuSingle.cpp:655,837).
#include "b2s_internal.h"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>

int b2s_fail(b2s_ctx* ctx, int code, const char* fmt, ...){
  char buf[1024];
  va_list ap; va_start(ap,fmt); vsnprintf(buf,sizeof(buf),fmt,ap); va_end(ap);
  if(ctx)ctx->err=buf;
  return code;
}

int b2s_set_dt_device(b2s_ctx* ctx, int which);
int b2s_set_dt_host(b2s_ctx* ctx, double dt);
int b2s_forces_fluid_only(b2s_ctx* ctx);

static void make_devcte(const b2s_params& p, DevCte& c){
  memset(&c,0,sizeof(c));
  c.kernelh=p.kernelh; c.ovkernelh=1.0f/p.kernelh; c.kernelsize2=p.kernelsize2; c.awen=p.awen; c.bwenh=p.bwenh;
  c.cub_a2=p.cubic_a2; c.cub_a24=p.cubic_a24; c.cub_c1=p.cubic_c1; c.cub_d1=p.cubic_d1; c.cub_c2=p.cubic_c2; c.cub_od_wdeltap=p.cubic_od_wdeltap;
  c.cteb=p.cteb; c.gamma=p.gamma; c.ovgamma=1.f/p.gamma; c.rhopzero=p.rhopzero; c.ovrhopzero=p.ovrhopzero;
  c.cs0=float(p.cs0); c.eta2=p.eta2; c.massf=p.massfluid; c.massb=p.massbound;
  c.ddtkh=p.ddtkh; c.ddtgz=p.ddtgz; c.spssmag=p.spssmag; c.spsblin=p.spsblin;
  c.scell=p.scell; c.movlimit=p.movlimit; c.rhopoutmin=p.rhopoutmin; c.rhopoutmax=p.rhopoutmax;
  c.visco_f=p.visco; c.visco_b=p.visco*p.viscoboundfactor;
  for(int k=0;k<3;k++){ c.gravity[k]=p.gravity[k]; c.domposmin[k]=p.domposmin[k]; c.maprealposmin[k]=p.maprealposmin[k]; c.maprealsize[k]=p.maprealsize[k]; c.domcells[k]=p.domcells[k]; }
  // (1+x)^a - 1 = a x + a(a-1)/2 x^2 + a(a-1)(a-2)/6 x^3,  a = 1/gamma   (DDT2 hydrostatic term)
  const double a=1.0/double(p.gamma);
  c.ddt_c1=float(a); c.ddt_c2=float(a*(a-1.)/2.); c.ddt_c3=float(a*(a-1.)*(a-2.)/6.);
  c.scell_d=double(p.scell); c.scell_blk=p.scell*float(1<<B2S_XBLK);
  c.cs0_d=p.cs0; c.kernelh_d=double(p.kernelh); c.cfl=p.cflnumber; c.dtmin=p.dtmin;
  c.domcellcode=p.domcellcode; c.scelldiv=p.scelldiv; c.tkernel=p.tkernel; c.tvisco=p.tvisco; c.tdensity=p.tdensity;
  c.tboundary=p.tboundary; c.tstep=p.tstep; c.use_floating=p.use_floating; c.dt_allparticles=p.dt_allparticles;
  c.refsem=p.refsem; c.simulate2d=p.simulate2d;
  c.shiftmode=p.shiftmode; c.shiftcoef=p.shiftcoef; c.shifttfs=p.shifttfs; c.shiftmaxdist=float(p.dp*0.1);
  c.mdbc2=(p.tboundary==B2S_BC_MDBC && p.slipmode>=B2S_SLIP_NOSLIP)? 1: 0; c.dp2=float(p.dp*p.dp);
  c.gamma_is7=(p.gamma==7.0f? 1: 0);
  c.periactive=p.periactive;
  for(int k=0;k<3;k++){ c.perinc[0][k]=p.perixinc[k]; c.perinc[1][k]=p.periyinc[k]; c.perinc[2][k]=p.perizinc[k]; }
}

template<class T> static cudaError_t dalloc(T*& ptr, size_t n){ ptr=NULL; return cudaMalloc((void**)&ptr, std::max<size_t>(n,1)*sizeof(T)); }

static cudaError_t alloc_partset(PartSet& s, size_t cap){
  cudaError_t e;
  if((e=dalloc(s.idp,cap)))return e;   if((e=dalloc(s.code,cap)))return e;  if((e=dalloc(s.dcell,cap)))return e;
  if((e=dalloc(s.posxy,cap)))return e; if((e=dalloc(s.posz,cap)))return e;  if((e=dalloc(s.velrho,cap)))return e;
  return cudaSuccess;
}
static void free_partset(PartSet& s){ cudaFree(s.idp); cudaFree(s.code); cudaFree(s.dcell); cudaFree(s.posxy); cudaFree(s.posz); cudaFree(s.velrho); memset(&s,0,sizeof(s)); }

extern "C" int b2s_create(const b2s_params* p, int device, uint32_t capacity_np, b2s_ctx** out){
  if(!p || !out)return B2S_ERR_ARG;
  *out=NULL;
  int ndev=0;
  if(cudaGetDeviceCount(&ndev)!=cudaSuccess || ndev<=0)return B2S_ERR_NODEVICE;
  if(device>=ndev)return B2S_ERR_ARG;
  if(p->periactive && p->casenfloat)return B2S_ERR_UNSUPPORTED;   // periodic + floating bodies: not built
  b2s_ctx* ctx=new b2s_ctx();
  ctx->prm=*p;
  if(ctx->prm.domcellcode==0){ int r=b2s_config_finalize(&ctx->prm); if(r){ delete ctx; return r; } }
  if(device<0)cudaGetDevice(&device);
  ctx->device=device;
  if(cudaSetDevice(device)!=cudaSuccess){ delete ctx; return B2S_ERR_NODEVICE; }
  make_devcte(ctx->prm,ctx->cte);
  ctx->cap=std::max<uint32_t>(capacity_np,1024u);
  ctx->np_host=ctx->npb_host=ctx->npbok_host=ctx->npf_host=0; ctx->slab_comm=NULL;
  ctx->launches=0; ctx->variant=5; ctx->pre_active=false; ctx->bound_changed=true; ctx->ftcount=0;
  ctx->verletstep=0;
  ctx->slab_on=false; ctx->slab_rank=0; ctx->slab_nranks=1; ctx->slab_cxa=0; ctx->slab_cxb=0; ctx->slab_counts=NULL; ctx->slab_flags=NULL; ctx->slab_counts_len=0;
  ctx->tile_bricks=NULL; ctx->tile_counters=NULL; ctx->tile_nsm=0; ctx->tile_cap=0;
  ctx->boundmode=NULL; ctx->tangenvel=NULL; ctx->shiftpos=NULL; ctx->damp_dev=NULL; ctx->damp_dirty=false; ctx->out_dev=NULL; ctx->out_count_dev=NULL; ctx->out_cap=0; ctx->motion_dev=NULL; ctx->motion_dirty=false; ctx->nearbound_valid=false;
  ctx->gauges_dev=NULL; ctx->gauge_res_dev=NULL; ctx->gauge_res_h=NULL; ctx->gauge_mass_dev=NULL; ctx->gauge_points=0; ctx->gauge_force_dev=NULL; ctx->gauge_force_n=0;
  ctx->ft_massp_dev=NULL; ctx->ft_sums_dev=NULL; ctx->ft_allsums_dev=NULL; ctx->ft_sums_h=NULL; ctx->ft_dev=NULL; ctx->scratch=NULL; ctx->scratch_bytes=0;
  ctx->pospre_xy=ctx->pospre_xy_alt=NULL; ctx->pospre_z=ctx->pospre_z_alt=NULL; ctx->velrhopre=ctx->velrhopre_alt=NULL;
  ctx->velrhom1=ctx->velrhom1_alt=NULL; ctx->spstau=ctx->spstau_alt=ctx->spstau_new=NULL; ctx->boundnor=ctx->boundnor_alt=NULL; ctx->motionace=ctx->motionace_alt=NULL;
  const size_t cap=ctx->cap;
  const b2s_params& q=ctx->prm;
  const uint64_t nct64=uint64_t(q.domcells[0])*q.domcells[1]*q.domcells[2];
  if(nct64>=0x7fffffffull/2){ delete ctx; return B2S_ERR_UNSUPPORTED; }
  ctx->nct_max=uint32_t(nct64); ctx->nctt_max=ctx->nct_max*2+6;
  cudaError_t e=cudaSuccess;
  #define TRY(x) do{ if(e==cudaSuccess)e=(x); }while(0)
  TRY(cudaStreamCreateWithFlags(&ctx->stream,cudaStreamNonBlocking));
  TRY(alloc_partset(ctx->cur,cap)); TRY(alloc_partset(ctx->alt,cap));
  if(q.tstep==B2S_STEP_SYMPLECTIC){
    TRY(dalloc(ctx->pospre_xy,cap)); TRY(dalloc(ctx->pospre_z,cap)); TRY(dalloc(ctx->velrhopre,cap));
    TRY(dalloc(ctx->pospre_xy_alt,cap)); TRY(dalloc(ctx->pospre_z_alt,cap)); TRY(dalloc(ctx->velrhopre_alt,cap));
  }
  else{ TRY(dalloc(ctx->velrhom1,cap)); TRY(dalloc(ctx->velrhom1_alt,cap)); }
  if(q.tvisco==B2S_VISCO_LAMINARSPS){
    TRY(dalloc(ctx->spstau,cap*6)); TRY(dalloc(ctx->spstau_alt,cap*6)); TRY(dalloc(ctx->spstau_new,cap*6));
    if(e==cudaSuccess){ cudaMemset(ctx->spstau,0,cap*24); cudaMemset(ctx->spstau_alt,0,cap*24); cudaMemset(ctx->spstau_new,0,cap*24); }
  }
  if(q.shiftmode){ TRY(dalloc(ctx->shiftpos,cap)); if(e==cudaSuccess)cudaMemset(ctx->shiftpos,0,size_t(cap)*16); }
  if(q.tboundary==B2S_BC_MDBC && q.slipmode>=B2S_SLIP_NOSLIP){ TRY(dalloc(ctx->boundmode,cap)); TRY(dalloc(ctx->tangenvel,cap)); if(e==cudaSuccess){ cudaMemset(ctx->boundmode,0,cap); cudaMemset(ctx->tangenvel,0,size_t(cap)*16); } }
  if(q.tboundary==B2S_BC_MDBC && q.slipmode>=B2S_SLIP_NOSLIP && q.casenmoving){ TRY(dalloc(ctx->motionace,cap)); TRY(dalloc(ctx->motionace_alt,cap)); if(e==cudaSuccess){ cudaMemset(ctx->motionace,0,size_t(cap)*16); cudaMemset(ctx->motionace_alt,0,size_t(cap)*16);} }
  if(q.tboundary==B2S_BC_MDBC){ TRY(dalloc(ctx->boundnor,cap)); TRY(dalloc(ctx->boundnor_alt,cap)); if(e==cudaSuccess){ cudaMemset(ctx->boundnor,0,cap*16); cudaMemset(ctx->boundnor_alt,0,cap*16);} }
  TRY(dalloc(ctx->posp,cap)); TRY(dalloc(ctx->acear,cap)); TRY(dalloc(ctx->posh,size_t(cap)/2+2+16));   // +16 pairs: the pre-filter may read one chunk past the last pair (forces_pair2.cu P2_CHK2)
  TRY(dalloc(ctx->cellpart,cap)); TRY(dalloc(ctx->sortpart,cap)); TRY(dalloc(ctx->cellpart_alt,cap)); TRY(dalloc(ctx->sortpart_alt,cap));
  { uint32_t* b=NULL; TRY(dalloc(b,size_t(ctx->nctt_max)+2)); ctx->begincell=(int2*)b; }
  TRY(dalloc(ctx->nearbound,size_t(ctx->nct_max)+8));
  ctx->radix_counts_len=size_t(512)*((cap+4095)/4096+1);   // digit histograms per tile: up to 9-bit digits (celldiv.cu radix_plan)
  TRY(dalloc(ctx->radix_counts,ctx->radix_counts_len+ctx->radix_counts_len/4096+8));   // + block totals of the scan (celldiv.cu)
  TRY(dalloc(ctx->ridp,size_t(q.casenmoving)+q.casenfloat+1));
  TRY(dalloc(ctx->div,1)); TRY(dalloc(ctx->step,1));
  TRY(cudaMallocHost((void**)&ctx->div_h,sizeof(DivState))); TRY(cudaMallocHost((void**)&ctx->step_h,sizeof(StepState)));
  #undef TRY
  if(e!=cudaSuccess){ b2s_destroy(ctx); return (e==cudaErrorMemoryAllocation? B2S_ERR_NOMEM: (int)e); }
  cudaMemset(ctx->div,0,sizeof(DivState)); cudaMemset(ctx->step,0,sizeof(StepState));
  memset(ctx->div_h,0,sizeof(DivState)); memset(ctx->step_h,0,sizeof(StepState));
  cudaMemset(ctx->acear,0,cap*sizeof(float4));
  *out=ctx;
  return B2S_OK;
}

extern "C" int b2s_destroy(b2s_ctx* ctx){
  if(!ctx)return B2S_OK;
  cudaSetDevice(ctx->device);
  if(ctx->stream)cudaStreamSynchronize(ctx->stream);
  b2s_slab_comm_free(ctx);
  free_partset(ctx->cur); free_partset(ctx->alt);
  cudaFree(ctx->pospre_xy); cudaFree(ctx->pospre_z); cudaFree(ctx->velrhopre);
  cudaFree(ctx->pospre_xy_alt); cudaFree(ctx->pospre_z_alt); cudaFree(ctx->velrhopre_alt);
  cudaFree(ctx->velrhom1); cudaFree(ctx->velrhom1_alt);
  cudaFree(ctx->spstau); cudaFree(ctx->spstau_alt); cudaFree(ctx->spstau_new);
  cudaFree(ctx->boundnor); cudaFree(ctx->boundnor_alt); cudaFree(ctx->motionace); cudaFree(ctx->motionace_alt);
  cudaFree(ctx->posp); cudaFree(ctx->acear); cudaFree(ctx->posh); cudaFree(ctx->tile_bricks); cudaFree(ctx->tile_counters);
  cudaFree(ctx->cellpart); cudaFree(ctx->sortpart); cudaFree(ctx->cellpart_alt); cudaFree(ctx->sortpart_alt);
  cudaFree(ctx->begincell); cudaFree(ctx->nearbound); cudaFree(ctx->radix_counts); cudaFree(ctx->ridp);
  cudaFree(ctx->div); cudaFree(ctx->step);
  cudaFree(ctx->gauges_dev); cudaFree(ctx->gauge_res_dev); cudaFree(ctx->gauge_force_dev);   // gauge_mass_dev lives inside gauge_res_dev
  if(ctx->gauge_res_h)cudaFreeHost(ctx->gauge_res_h);
  cudaFree(ctx->motion_dev); cudaFree(ctx->out_dev); cudaFree(ctx->out_count_dev); cudaFree(ctx->damp_dev); cudaFree(ctx->shiftpos); cudaFree(ctx->boundmode); cudaFree(ctx->tangenvel); cudaFree(ctx->slab_counts); cudaFree(ctx->slab_flags);
  cudaFree(ctx->ft_massp_dev); cudaFree(ctx->ft_sums_dev); cudaFree(ctx->ft_allsums_dev); cudaFree(ctx->ft_dev); cudaFree(ctx->scratch);
  if(ctx->ft_sums_h)cudaFreeHost(ctx->ft_sums_h);
  if(ctx->div_h)cudaFreeHost(ctx->div_h);
  if(ctx->step_h)cudaFreeHost(ctx->step_h);
  if(ctx->stream)cudaStreamDestroy(ctx->stream);
  delete ctx;
  return B2S_OK;
}

extern "C" const char* b2s_last_error(const b2s_ctx* ctx){ return ctx? ctx->err.c_str(): "null context"; }
extern "C" int b2s_get_params(const b2s_ctx* ctx, b2s_params* out){ if(!ctx||!out)return B2S_ERR_ARG; *out=ctx->prm; return B2S_OK; }
extern "C" int b2s_get_stream(b2s_ctx* ctx, void** s){ if(!ctx||!s)return B2S_ERR_ARG; *s=(void*)ctx->stream; return B2S_OK; }
extern "C" int b2s_launch_count(b2s_ctx* ctx, uint64_t* l, int reset){ if(!ctx)return B2S_ERR_ARG; if(l)*l=ctx->launches; if(reset)ctx->launches=0; return B2S_OK; }
extern "C" int b2s_set_interaction_variant(b2s_ctx* ctx, int v){ if(!ctx||(v!=0&&v!=5&&v!=6))return B2S_ERR_ARG; ctx->variant=v; return B2S_OK; }
// Test hook: lowers the particle capacity of a shared-memory tile of interaction variant 6 (0 = default) so that small cases
// exercise the column split and the global-memory fallback.
extern "C" int b2s_debug_set_tile_cap(b2s_ctx* ctx, uint32_t cap){ if(!ctx)return B2S_ERR_ARG; ctx->tile_cap=cap; return B2S_OK; }

// Copies DivState+StepState to pinned memory, synchronises, refreshes the host mirrors and checks device flags.
int b2s_sync_info(b2s_ctx* ctx){
  B2S_CUDA(ctx,cudaMemcpyAsync(ctx->div_h,ctx->div,sizeof(DivState),cudaMemcpyDeviceToHost,ctx->stream));
  B2S_CUDA(ctx,cudaMemcpyAsync(ctx->step_h,ctx->step,sizeof(StepState),cudaMemcpyDeviceToHost,ctx->stream));
  B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  ctx->np_host=ctx->div_h->np; ctx->npb_host=ctx->div_h->npb; ctx->npbok_host=ctx->div_h->npbok; ctx->npf_host=ctx->np_host-ctx->npb_host;
  if(ctx->div_h->error==2u)return b2s_fail(ctx,B2S_ERR_NOMEM,"particle capacity %u exceeded by the periodic duplicates",ctx->cap);
  if(ctx->div_h->error)return b2s_fail(ctx,B2S_ERR_BOUNDOUT,"boundary particles were excluded from the domain (%u)",ctx->div_h->npb_out);
  if(ctx->step_h->error)return b2s_fail(ctx,ctx->step_h->error,"%s (acemax=%g velmax=%g viscdtmax=%g nstep=%u)",b2s_error_string(ctx->step_h->error),ctx->step_h->acemax,ctx->step_h->velmax,(double)ctx->step_h->viscdtmax,ctx->step_h->nstep);
  return B2S_OK;
}

// JSph::LoadDcellParticles (JSph.cpp:1794-1820) on the device.
__global__ void k_init_dcell(uint32_t n, const double2* __restrict__ posxy, const double* __restrict__ posz, const typecode* __restrict__ code,
  uint32_t* __restrict__ dcell, int* err, const __grid_constant__ DevCte c, double rminx,double rminy,double rminz,double rmaxx,double rmaxy,double rmaxz)
{
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=n)return;
  if(B2S_CODE_SPECIAL(code[p])<B2S_CODE_OUTIGNORE){
    const double2 xy=posxy[p]; const double z=posz[p];
    if(xy.x>=rminx && xy.y>=rminy && z>=rminz && xy.x<rmaxx && xy.y<rmaxy && z<rmaxz){
      const uint32_t cx=uint32_t((xy.x-c.domposmin[0])/c.scell_d), cy=uint32_t((xy.y-c.domposmin[1])/c.scell_d), cz=uint32_t((z-c.domposmin[2])/c.scell_d);
      dcell[p]=B2S_DCEL_CELL(c.domcellcode,cx,cy,cz);
    }
    else{ *err=1; dcell[p]=B2S_DCEL_MAPOUT; }
  }
  else dcell[p]=B2S_DCEL_MAPOUT;
}
__global__ void k_split_pos(uint32_t n, const double* __restrict__ pos3, double2* __restrict__ posxy, double* __restrict__ posz){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p<n){ posxy[p]=make_double2(pos3[size_t(p)*3],pos3[size_t(p)*3+1]); posz[p]=pos3[size_t(p)*3+2]; }
}
__global__ void k_join_pos(uint32_t n, const double2* __restrict__ posxy, const double* __restrict__ posz, double* __restrict__ pos3){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p<n){ const double2 xy=posxy[p]; pos3[size_t(p)*3]=xy.x; pos3[size_t(p)*3+1]=xy.y; pos3[size_t(p)*3+2]=posz[p]; }
}
__global__ void k_nor3to4(uint32_t n, const float* __restrict__ n3, float4* __restrict__ n4){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p<n)n4[p]=make_float4(n3[size_t(p)*3],n3[size_t(p)*3+1],n3[size_t(p)*3+2],0.f);
}
__global__ void k_init_state(DivState* ds, StepState* st, uint32_t np, uint32_t npb, double dtini){
  ds->np=np; ds->npb=npb; ds->npbok=npb; ds->nptot=np; ds->npf_out_total=0; ds->error=0;
  memset(st,0,sizeof(StepState));
  st->symplectic_dtpre=dtini; st->dt=dtini;
}

static int ensure_scratch(b2s_ctx* ctx, size_t bytes){
  if(ctx->scratch_bytes>=bytes)return B2S_OK;
  cudaFree(ctx->scratch); ctx->scratch=NULL; ctx->scratch_bytes=0;
  B2S_CUDA(ctx,cudaMalloc(&ctx->scratch,bytes));
  ctx->scratch_bytes=bytes;
  return B2S_OK;
}

extern "C" int b2s_upload_particles(b2s_ctx* ctx, uint32_t np, const uint32_t* idp, const uint16_t* code,
                         const double* pos, const float* velrho, const float* boundnor)
{
  if(!ctx || !idp || !code || !pos || !velrho)return B2S_ERR_ARG;
  if(np>ctx->cap)return b2s_fail(ctx,B2S_ERR_ARG,"np=%u exceeds capacity %u",np,ctx->cap);
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  cudaStream_t s=ctx->stream;
  int r=ensure_scratch(ctx,size_t(np)*24+16); if(r)return r;
  B2S_CUDA(ctx,cudaMemcpyAsync(ctx->cur.idp,idp,size_t(np)*4,cudaMemcpyHostToDevice,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(ctx->cur.code,code,size_t(np)*2,cudaMemcpyHostToDevice,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(ctx->cur.velrho,velrho,size_t(np)*16,cudaMemcpyHostToDevice,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(ctx->scratch,pos,size_t(np)*24,cudaMemcpyHostToDevice,s));
  k_split_pos<<<b2s_blocks(np,256),256,0,s>>>(np,(const double*)ctx->scratch,ctx->cur.posxy,ctx->cur.posz); B2S_LAUNCHED(ctx);
  if(ctx->boundnor){
    B2S_CUDA(ctx,cudaMemsetAsync(ctx->boundnor,0,size_t(ctx->cap)*16,s));
    if(boundnor){
      // normals are given for the casenbound first particles; the solver doubles them to reach the ghost node (JSph.cpp:1417)
      // (the count of leading non-fluid particles, by bisection: a slab upload holds a subset of the case's boundary)
      uint32_t nb=0;
      { uint32_t lo=0,hi=np; while(lo<hi){ const uint32_t mid=lo+(hi-lo)/2; if(!B2S_CODE_ISFLUID(code[mid]))lo=mid+1; else hi=mid; } nb=lo; }
      B2S_CUDA(ctx,cudaStreamSynchronize(s));
      B2S_CUDA(ctx,cudaMemcpyAsync(ctx->scratch,boundnor,size_t(nb)*12,cudaMemcpyHostToDevice,s));
      k_nor3to4<<<b2s_blocks(nb,256),256,0,s>>>(nb,(const float*)ctx->scratch,ctx->boundnor); B2S_LAUNCHED(ctx);
    }
  }
  // npb = bound particles = non-floating boundary (fixed+moving), they come first
  // (found by bisection on the type bits: callers may upload a subset of the case, e.g. one slab, so CaseNpb is only a bound)
  uint32_t npb=0;
  { uint32_t lo=0,hi=np; while(lo<hi){ const uint32_t mid=lo+(hi-lo)/2; if(B2S_CODE_TYPE(code[mid])<B2S_CODE_TYPE_FLOATING)lo=mid+1; else hi=mid; } npb=lo; }
  if(npb>ctx->prm.casenpb)return b2s_fail(ctx,B2S_ERR_ARG,"%u boundary particles uploaded but the case has %u",npb,ctx->prm.casenpb);
  int* derr=(int*)((char*)ctx->scratch+ctx->scratch_bytes-8);
  B2S_CUDA(ctx,cudaStreamSynchronize(s));
  B2S_CUDA(ctx,cudaMemsetAsync(derr,0,4,s));
  const b2s_params& q=ctx->prm;
  k_init_dcell<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->cur.posxy,ctx->cur.posz,ctx->cur.code,ctx->cur.dcell,derr,ctx->cte,
    q.domrealposmin[0],q.domrealposmin[1],q.domrealposmin[2],q.domrealposmax[0],q.domrealposmax[1],q.domrealposmax[2]); B2S_LAUNCHED(ctx);
  int herr=0;
  B2S_CUDA(ctx,cudaMemcpyAsync(&herr,derr,4,cudaMemcpyDeviceToHost,s));
  k_init_state<<<1,1,0,s>>>(ctx->div,ctx->step,np,npb,q.dtini); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaStreamSynchronize(s));
  if(herr)return b2s_fail(ctx,B2S_ERR_ARG,"found initial particles outside the domain (JSph::LoadDcellParticles)");
  ctx->np_host=np; ctx->npb_host=npb; ctx->npbok_host=npb; ctx->npf_host=np-npb;
  ctx->pre_active=false; ctx->bound_changed=true; ctx->verletstep=0;
  if(ctx->velrhom1)B2S_CUDA(ctx,cudaMemcpyAsync(ctx->velrhom1,ctx->cur.velrho,size_t(np)*16,cudaMemcpyDeviceToDevice,s));  // InitRunGpu: VelrhoM1=Velrho
  if(ctx->motionace)B2S_CUDA(ctx,cudaMemsetAsync(ctx->motionace,0,size_t(ctx->cap)*16,s));
  r=b2s_periodic_impl(ctx); if(r)return r;
  r=b2s_celldiv_impl(ctx); if(r)return r;
  return b2s_sync_info(ctx);
}

__global__ void k_set_time(StepState* st, double timestep, double dtpre){ st->timestep=timestep; if(dtpre>0){ st->symplectic_dtpre=dtpre; st->dt=dtpre; } }
// Restart: simulated time and (Symplectic) the dt of the next step, as stored in the PART the run continues from.
extern "C" int b2s_set_time(b2s_ctx* ctx, double timestep, double symplectic_dtpre){
  if(!ctx || !(timestep>=0))return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  k_set_time<<<1,1,0,ctx->stream>>>(ctx->step,timestep,symplectic_dtpre); B2S_LAUNCHED(ctx);
  return b2s_sync_info(ctx);
}

extern "C" int b2s_download_particles(b2s_ctx* ctx, uint32_t* np_out, uint32_t* idp, uint16_t* code, double* pos, float* velrho){
  if(!ctx)return B2S_ERR_ARG;
  int r=b2s_sync_info(ctx); if(r)return r;
  const uint32_t np=ctx->np_host;
  cudaStream_t s=ctx->stream;
  if(np_out)*np_out=np;
  if(idp)B2S_CUDA(ctx,cudaMemcpyAsync(idp,ctx->cur.idp,size_t(np)*4,cudaMemcpyDeviceToHost,s));
  if(code)B2S_CUDA(ctx,cudaMemcpyAsync(code,ctx->cur.code,size_t(np)*2,cudaMemcpyDeviceToHost,s));
  if(velrho)B2S_CUDA(ctx,cudaMemcpyAsync(velrho,ctx->cur.velrho,size_t(np)*16,cudaMemcpyDeviceToHost,s));
  if(pos){
    r=ensure_scratch(ctx,size_t(np)*24+16); if(r)return r;
    k_join_pos<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->cur.posxy,ctx->cur.posz,(double*)ctx->scratch); B2S_LAUNCHED(ctx);
    B2S_CUDA(ctx,cudaMemcpyAsync(pos,ctx->scratch,size_t(np)*24,cudaMemcpyDeviceToHost,s));
  }
  B2S_CUDA(ctx,cudaStreamSynchronize(s));
  return B2S_OK;
}

// scatter by id: out[idp]=value, compacted afterwards on the host side (ids are dense while no particle is out).
__global__ void k_scatter_byid(uint32_t n, const uint32_t* __restrict__ idp, const typecode* __restrict__ code, const double2* __restrict__ posxy, const double* __restrict__ posz,
  const float4* __restrict__ velrho, uint32_t idmax, double* pos3, float* vel3, float* rhop, uint32_t* present)
{
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=n)return;
  const uint32_t id=idp[p];
  if(id>=idmax || !B2S_CODE_ISNORMAL(code[p]))return;   // periodic duplicates and slab ghosts share the id of their original
  const double2 xy=posxy[p]; const float4 v=velrho[p];
  pos3[size_t(id)*3]=xy.x; pos3[size_t(id)*3+1]=xy.y; pos3[size_t(id)*3+2]=posz[p];
  vel3[size_t(id)*3]=v.x; vel3[size_t(id)*3+1]=v.y; vel3[size_t(id)*3+2]=v.z; rhop[id]=v.w; present[id]=1u;
}

extern "C" int b2s_download_particles_byid(b2s_ctx* ctx, uint32_t* np_out, uint32_t* idp, double* pos, float* vel, float* rhop){
  if(!ctx || !pos || !vel || !rhop || !idp)return B2S_ERR_ARG;
  int r=b2s_sync_info(ctx); if(r)return r;
  const uint32_t np=ctx->np_host, idmax=ctx->prm.casenp;
  cudaStream_t s=ctx->stream;
  if(ctx->slab_on && ctx->slab_nranks>1){
    // A slab rank owns 1/N of the case: everything is sized by what it holds (np), not by the case (casenp) -- the particle arrays
    // come down as they are and the owned particles are ordered by id on the host.
    std::vector<uint32_t> hid(np); std::vector<typecode> hcode(np); std::vector<double2> hxy(np); std::vector<double> hz(np); std::vector<float4> hvr(np);
    B2S_CUDA(ctx,cudaMemcpyAsync(hid.data(),ctx->cur.idp,size_t(np)*4,cudaMemcpyDeviceToHost,s));
    B2S_CUDA(ctx,cudaMemcpyAsync(hcode.data(),ctx->cur.code,size_t(np)*sizeof(typecode),cudaMemcpyDeviceToHost,s));
    B2S_CUDA(ctx,cudaMemcpyAsync(hxy.data(),ctx->cur.posxy,size_t(np)*16,cudaMemcpyDeviceToHost,s));
    B2S_CUDA(ctx,cudaMemcpyAsync(hz.data(),ctx->cur.posz,size_t(np)*8,cudaMemcpyDeviceToHost,s));
    B2S_CUDA(ctx,cudaMemcpyAsync(hvr.data(),ctx->cur.velrho,size_t(np)*16,cudaMemcpyDeviceToHost,s));
    B2S_CUDA(ctx,cudaStreamSynchronize(s));
    std::vector<uint32_t> order; order.reserve(np);
    for(uint32_t p=0;p<np;p++)if(B2S_CODE_ISNORMAL(hcode[p]) && hid[p]<idmax)order.push_back(p);   // ghosts and excluded particles stay out (k_scatter_byid's rule)
    std::sort(order.begin(),order.end(),[&](uint32_t a,uint32_t b){ return hid[a]<hid[b]; });
    uint32_t n=0;
    for(uint32_t p:order){
      idp[n]=hid[p]; pos[size_t(n)*3]=hxy[p].x; pos[size_t(n)*3+1]=hxy[p].y; pos[size_t(n)*3+2]=hz[p];
      vel[size_t(n)*3]=hvr[p].x; vel[size_t(n)*3+1]=hvr[p].y; vel[size_t(n)*3+2]=hvr[p].z; rhop[n]=hvr[p].w; n++;
    }
    if(np_out)*np_out=n;
    return B2S_OK;
  }
  const size_t bytes=size_t(idmax)*(24+12+4+4)+64;
  r=ensure_scratch(ctx,bytes); if(r)return r;
  char* base=(char*)ctx->scratch;
  double* dpos=(double*)base; float* dvel=(float*)(base+size_t(idmax)*24); float* drho=(float*)(base+size_t(idmax)*36); uint32_t* dpres=(uint32_t*)(base+size_t(idmax)*40);
  B2S_CUDA(ctx,cudaMemsetAsync(dpres,0,size_t(idmax)*4,s));
  k_scatter_byid<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->cur.idp,ctx->cur.code,ctx->cur.posxy,ctx->cur.posz,ctx->cur.velrho,idmax,dpos,dvel,drho,dpres); B2S_LAUNCHED(ctx);
  std::vector<uint32_t> pres(idmax);
  std::vector<double> hpos(size_t(idmax)*3); std::vector<float> hvel(size_t(idmax)*3), hrho(idmax);
  B2S_CUDA(ctx,cudaMemcpyAsync(pres.data(),dpres,size_t(idmax)*4,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(hpos.data(),dpos,size_t(idmax)*24,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(hvel.data(),dvel,size_t(idmax)*12,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(hrho.data(),drho,size_t(idmax)*4,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaStreamSynchronize(s));
  uint32_t n=0;
  for(uint32_t id=0;id<idmax;id++)if(pres[id]){
    idp[n]=id; pos[size_t(n)*3]=hpos[size_t(id)*3]; pos[size_t(n)*3+1]=hpos[size_t(id)*3+1]; pos[size_t(n)*3+2]=hpos[size_t(id)*3+2];
    vel[size_t(n)*3]=hvel[size_t(id)*3]; vel[size_t(n)*3+1]=hvel[size_t(id)*3+1]; vel[size_t(n)*3+2]=hvel[size_t(id)*3+2]; rhop[n]=hrho[id]; n++;
  }
  if(np_out)*np_out=n;
  return B2S_OK;
}

__global__ void k_poscell_ref(uint32_t n, const double2* __restrict__ posxy, const double* __restrict__ posz, float4* out, double mx,double my,double mz, float pcs){
  // cusphs::KerUpdatePosCell (JSphGpuSimple_ker.cu:41-69), reference layout 13/10/9 bits.
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=n)return;
  const double2 xy=posxy[p];
  const double dx=xy.x-mx, dy=xy.y-my, dz=posz[p]-mz;
  const uint32_t cx=uint32_t(dx/pcs), cy=uint32_t(dy/pcs), cz=uint32_t(dz/pcs);
  const float px=float(dx-(double(pcs)*cx)), py=float(dy-(double(pcs)*cy)), pz=float(dz-(double(pcs)*cz));
  const uint32_t w=(cx<<19)|((cy<<9)&0x0007fe00u)|(cz&0x000001ffu);
  out[p]=make_float4(px,py,pz,__uint_as_float(w));
}
__global__ void k_unpack4(uint32_t n, const float4* __restrict__ src, float* xyz, float* w){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=n)return;
  const float4 v=src[p];
  if(xyz){ xyz[size_t(p)*3]=v.x; xyz[size_t(p)*3+1]=v.y; xyz[size_t(p)*3+2]=v.z; }
  if(w)w[p]=v.w;
}
__global__ void k_beginend_int2(uint32_t nboxes, const uint32_t* __restrict__ beg, int2* out){
  const uint32_t c=blockIdx.x*blockDim.x+threadIdx.x;
  if(c>=nboxes)return;
  const uint32_t a=beg[c], b=beg[c+1];
  out[c]=(a==b? make_int2(0,0): make_int2(int(a),int(b)));
}

// ---- positions by id through ridp (include/b2s.h) ---------------------------------------------------------------------------
__global__ void k_pos_byid(uint32_t n,const uint32_t* __restrict__ ids,const uint32_t* __restrict__ ridp,uint32_t nfixed,uint32_t nmot,uint32_t np,
  const uint32_t* __restrict__ idp,const double2* __restrict__ posxy,const double* __restrict__ posz,double* __restrict__ out)
{
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=n)return;
  const double nan=__longlong_as_double(0x7ff8000000000000ll);
  double x=nan,y=nan,z=nan;
  const uint32_t id=ids[i];
  if(id>=nfixed && id-nfixed<nmot){
    const uint32_t p=ridp[id-nfixed];
    if(p<np && idp[p]==id){ const double2 xy=posxy[p]; x=xy.x; y=xy.y; z=posz[p]; }   // a stale entry (particle excluded since) fails the id test
  }
  out[size_t(i)*3]=x; out[size_t(i)*3+1]=y; out[size_t(i)*3+2]=z;
}
extern "C" int b2s_get_pos_byid(b2s_ctx* ctx, uint32_t n, const uint32_t* ids, double* pos_out){
  if(!ctx || (n && (!ids || !pos_out)))return B2S_ERR_ARG;
  if(!n)return B2S_OK;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  int r=b2s_sync_info(ctx); if(r)return r;
  r=ensure_scratch(ctx,size_t(n)*(4+24)+64); if(r)return r;
  cudaStream_t s=ctx->stream;
  uint32_t* dids=(uint32_t*)ctx->scratch; double* dout=(double*)((char*)ctx->scratch+((size_t(n)*4+63)/64)*64);
  B2S_CUDA(ctx,cudaMemcpyAsync(dids,ids,size_t(n)*4,cudaMemcpyHostToDevice,s));
  k_pos_byid<<<b2s_blocks(n,128),128,0,s>>>(n,dids,ctx->ridp,ctx->prm.casenfixed,ctx->prm.casenmoving+ctx->prm.casenfloat,ctx->np_host,
                                           ctx->cur.idp,ctx->cur.posxy,ctx->cur.posz,dout); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaMemcpyAsync(pos_out,dout,size_t(n)*24,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaStreamSynchronize(s));
  return B2S_OK;
}

extern "C" int b2s_get_array(b2s_ctx* ctx, const char* name, void* dst, size_t capacity_bytes, size_t* count_out){
  if(!ctx || !name || !dst)return B2S_ERR_ARG;
  int r=b2s_sync_info(ctx); if(r)return r;
  cudaStream_t s=ctx->stream;
  const uint32_t np=ctx->np_host;
  const std::string nm=name;
  const void* src=NULL; size_t esize=0, count=0;
  #define DIRECT(ptr,es,cnt) do{ src=(ptr); esize=(es); count=(cnt); }while(0)
  if(nm=="idp")DIRECT(ctx->cur.idp,4,np);
  else if(nm=="code")DIRECT(ctx->cur.code,2,np);
  else if(nm=="dcell")DIRECT(ctx->cur.dcell,4,np);
  else if(nm=="velrho")DIRECT(ctx->cur.velrho,16,np);
  else if(nm=="velrhopre" && ctx->velrhopre)DIRECT(ctx->velrhopre,16,np);
  else if(nm=="velrhom1" && ctx->velrhom1)DIRECT(ctx->velrhom1,16,np);
  else if(nm=="cellpart")DIRECT(ctx->cellpart,4,ctx->div_h->nptot);
  else if(nm=="sortpart")DIRECT(ctx->sortpart,4,ctx->div_h->nptot);
  else if(nm=="begincell")DIRECT(ctx->begincell,4,ctx->div_h->nctt);
  else if(nm=="spstau" && ctx->spstau)DIRECT(ctx->spstau,24,np);
  else if(nm=="posp")DIRECT(ctx->posp,16,np);
  else if(nm=="acear")DIRECT(ctx->acear,16,np);
  else if(nm=="boundnor" && ctx->boundnor)DIRECT(ctx->boundnor,16,np);
  else if(nm=="ridp")DIRECT(ctx->ridp,4,ctx->prm.casenmoving+ctx->prm.casenfloat);
  #undef DIRECT
  if(src){
    if(count*esize>capacity_bytes)return b2s_fail(ctx,B2S_ERR_ARG,"buffer too small for %s: need %zu bytes",name,count*esize);
    B2S_CUDA(ctx,cudaMemcpyAsync(dst,src,count*esize,cudaMemcpyDeviceToHost,s));
    B2S_CUDA(ctx,cudaStreamSynchronize(s));
    if(count_out)*count_out=count;
    return B2S_OK;
  }
  // derived views
  size_t bytes=0;
  if(nm=="pos" || nm=="pospre"){
    if(nm=="pospre" && !ctx->pospre_xy)return b2s_fail(ctx,B2S_ERR_NAME,"array %s not available",name);
    bytes=size_t(np)*24; r=ensure_scratch(ctx,bytes+16); if(r)return r;
    if(bytes>capacity_bytes)return b2s_fail(ctx,B2S_ERR_ARG,"buffer too small for %s",name);
    if(nm=="pos")k_join_pos<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->cur.posxy,ctx->cur.posz,(double*)ctx->scratch);
    else k_join_pos<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->pospre_xy,ctx->pospre_z,(double*)ctx->scratch);
    B2S_LAUNCHED(ctx); count=np;
  }
  else if(nm=="ace" || nm=="ar" || nm=="press"){
    bytes=size_t(np)*(nm=="ace"? 12: 4); r=ensure_scratch(ctx,bytes+16); if(r)return r;
    if(bytes>capacity_bytes)return b2s_fail(ctx,B2S_ERR_ARG,"buffer too small for %s",name);
    if(nm=="ace")k_unpack4<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->acear,(float*)ctx->scratch,NULL);
    else if(nm=="ar")k_unpack4<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->acear,NULL,(float*)ctx->scratch);
    else k_unpack4<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->posp,NULL,(float*)ctx->scratch);
    B2S_LAUNCHED(ctx); count=np;
  }
  else if(nm=="poscell"){
    bytes=size_t(np)*16; r=ensure_scratch(ctx,bytes+16); if(r)return r;
    if(bytes>capacity_bytes)return b2s_fail(ctx,B2S_ERR_ARG,"buffer too small for %s",name);
    k_poscell_ref<<<b2s_blocks(np,256),256,0,s>>>(np,ctx->cur.posxy,ctx->cur.posz,(float4*)ctx->scratch,ctx->prm.map_posmin[0],ctx->prm.map_posmin[1],ctx->prm.map_posmin[2],ctx->prm.poscellsize);
    B2S_LAUNCHED(ctx); count=np;
  }
  else if(nm=="beginendcell"){
    const uint32_t nb=ctx->div_h->nct*2+5;
    bytes=size_t(nb)*8; r=ensure_scratch(ctx,bytes+16); if(r)return r;
    if(bytes>capacity_bytes)return b2s_fail(ctx,B2S_ERR_ARG,"buffer too small for %s",name);
    k_beginend_int2<<<b2s_blocks(nb,256),256,0,s>>>(nb,(const uint32_t*)ctx->begincell,(int2*)ctx->scratch);
    B2S_LAUNCHED(ctx); count=nb;
  }
  else return b2s_fail(ctx,B2S_ERR_NAME,"unknown or unavailable array '%s'",name);
  B2S_CUDA(ctx,cudaMemcpyAsync(dst,ctx->scratch,bytes,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaStreamSynchronize(s));
  if(count_out)*count_out=count;
  return B2S_OK;
}

extern "C" int b2s_get_info(b2s_ctx* ctx, b2s_info* o){
  if(!ctx || !o)return B2S_ERR_ARG;
  int r=b2s_sync_info(ctx);
  const DivState& d=*ctx->div_h; const StepState& t=*ctx->step_h;
  memset(o,0,sizeof(*o));
  o->np=d.np; o->npb=d.npb; o->npbok=d.npbok; o->npf_out_total=d.npf_out_total;
  o->ncx=d.ncx; o->ncy=d.ncy; o->ncz=d.ncz; o->nct=d.nct;
  for(int k=0;k<3;k++){ o->cellmin[k]=d.cellmin[k]; o->cellmax[k]=d.cellmax[k]; }
  o->npb_ignore=d.npb_ignore; o->npf_out_last=d.npf_out; o->npb_out_last=d.npb_out; o->nstep=t.nstep;
  o->timestep=t.timestep; o->dt_last=t.dt_last; o->dt_p=t.dt_p; o->dt_c=t.dt_c; o->symplectic_dtpre=t.symplectic_dtpre;
  o->velmax=t.velmax; o->acemax=t.acemax; o->viscdtmax=t.viscdtmax; o->verletstep=ctx->verletstep; o->dtmodif=t.dtmodif;
  o->pairs_checked=t.pairs_checked; o->pairs_real=t.pairs_real;
  return r;
}

// ---- fine-grained entry points ----------------------------------------------------------------------
extern "C" int b2s_celldiv(b2s_ctx* ctx){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  int r=b2s_periodic_impl(ctx); if(r)return r;
  r=b2s_celldiv_impl(ctx); if(r)return r;
  return b2s_sync_info(ctx);
}
extern "C" int b2s_interaction(b2s_ctx* ctx, int interstep){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  if(ctx->prm.tboundary==B2S_BC_MDBC && (ctx->prm.mdbc_corrector || interstep!=B2S_INTER_SYMCORRECTOR)){ int r=b2s_mdbc_impl(ctx); if(r)return r; }
  int r=b2s_forces_impl(ctx,interstep); if(r)return r;
  return b2s_sync_info(ctx);
}
extern "C" int b2s_compute_verlet(b2s_ctx* ctx, double dt){ if(!ctx)return B2S_ERR_ARG; B2S_CUDA(ctx,cudaSetDevice(ctx->device)); return b2s_integrate_impl(ctx,0,dt,false); }
extern "C" int b2s_compute_symplectic_pre(b2s_ctx* ctx, double dt){ if(!ctx)return B2S_ERR_ARG; B2S_CUDA(ctx,cudaSetDevice(ctx->device)); return b2s_integrate_impl(ctx,1,dt,false); }
extern "C" int b2s_compute_symplectic_corr(b2s_ctx* ctx, double dt){ if(!ctx)return B2S_ERR_ARG; B2S_CUDA(ctx,cudaSetDevice(ctx->device)); return b2s_integrate_impl(ctx,2,dt,false); }

// ---- whole step -----------------------------------------------------------------------------------------
__global__ void k_step_end(StepState* st, int symplectic){
  if(symplectic)st->symplectic_dtpre=fmin(st->dt_p,st->dt_c);
  st->timestep+=st->dt;
  st->nstep++;
}

// RunCellDivide(true): periodic duplicates, then the divide.  Periodic runs refresh the host mirrors here (the number of
// duplicates decides the launch bounds); all other runs keep the one synchronisation per step.
static int divide_full(b2s_ctx* ctx){
  int r=b2s_periodic_impl(ctx); if(r)return r;
  r=b2s_celldiv_impl(ctx); if(r)return r;
  if(ctx->cte.periactive)r=b2s_sync_info(ctx);
  return r;
}

static int enqueue_step(b2s_ctx* ctx){
  int r;
  const b2s_params& q=ctx->prm;
  const bool mdbc=(q.tboundary==B2S_BC_MDBC);
  if(q.tstep==B2S_STEP_SYMPLECTIC){
    // ComputeStep_Sym (JSphGpuSingle.cpp:785-819)
    b2s_set_dt_device(ctx,0);                                   // dt=SymplecticDtPre
    if(mdbc){ r=b2s_mdbc_impl(ctx); if(r)return r; }
    r=b2s_forces_impl(ctx,B2S_INTER_SYMPREDICTOR); if(r)return r;  // -> dt_p
    r=b2s_integrate_impl(ctx,1,0,true); if(r)return r;
    if(ctx->ftcount){ r=b2s_floating_impl(ctx,0,true,true); if(r)return r; }
    r=divide_full(ctx); if(r)return r;
    if(mdbc && q.mdbc_corrector){ r=b2s_mdbc_impl(ctx); if(r)return r; }
    r=b2s_forces_impl(ctx,B2S_INTER_SYMCORRECTOR); if(r)return r;  // -> dt_c
    r=b2s_integrate_impl(ctx,2,0,true); if(r)return r;
    if(ctx->ftcount){ r=b2s_floating_impl(ctx,0,false,true); if(r)return r; }
  }
  else{
    // ComputeStep_Ver (JSphGpuSingle.cpp:761-776)
    if(mdbc){ r=b2s_mdbc_impl(ctx); if(r)return r; }
    r=b2s_forces_impl(ctx,B2S_INTER_VERLET); if(r)return r;        // -> dt_c
    b2s_set_dt_device(ctx,1);
    r=b2s_integrate_impl(ctx,0,0,true); if(r)return r;
    if(ctx->ftcount){ r=b2s_floating_impl(ctx,0,false,true); if(r)return r; }
  }
  r=b2s_damping_impl(ctx); if(r)return r;                           // RunDamping(dt), JSphCpuSingle.cpp:885,928
  // Run-loop tail (JSphGpuSingle.cpp:1003-1008): RunMotion, RunCellDivide, TimeStep+=dt
  // RunGaugeSystem(TimeStep+stepdt) (JSphCpuSingle.cpp:1192): new particle state, cell table of the last divide
  r=b2s_gauges_run(ctx,0); if(r)return r;
  if(!ctx->motions.empty()){ r=b2s_motion_impl(ctx,0,0); if(r)return r; }
  k_step_end<<<1,1,0,ctx->stream>>>(ctx->step,q.tstep==B2S_STEP_SYMPLECTIC? 1: 0); B2S_LAUNCHED(ctx);
  r=divide_full(ctx); if(r)return r;
  return B2S_OK;
}

extern "C" int b2s_compute_step(b2s_ctx* ctx, double* dt_out){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  int r=enqueue_step(ctx); if(r)return r;
  r=b2s_sync_info(ctx);
  if(dt_out)*dt_out=ctx->step_h->dt;
  return r;
}

extern "C" int b2s_run_steps(b2s_ctx* ctx, uint32_t nsteps, double* time_out){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  int r=B2S_OK;
  for(uint32_t i=0;i<nsteps && r==B2S_OK;i++){
    r=enqueue_step(ctx); if(r)break;
    r=b2s_sync_info(ctx);      // one synchronisation per step: refreshes np for the launch bounds, checks flags
  }
  if(time_out)*time_out=ctx->step_h->timestep;
  return r;
}

extern "C" int b2s_time_interaction(b2s_ctx* ctx, int reps, float* ms_per_launch){
  if(!ctx || reps<=0 || !ms_per_launch)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  cudaEvent_t e0,e1;
  B2S_CUDA(ctx,cudaEventCreate(&e0)); B2S_CUDA(ctx,cudaEventCreate(&e1));
  int r=b2s_forces_fluid_only(ctx); if(r)return r;   // warm-up
  B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  B2S_CUDA(ctx,cudaEventRecord(e0,ctx->stream));
  for(int i=0;i<reps;i++){ r=b2s_forces_fluid_only(ctx); if(r)return r; }
  B2S_CUDA(ctx,cudaEventRecord(e1,ctx->stream));
  B2S_CUDA(ctx,cudaEventSynchronize(e1));
  float ms=0; B2S_CUDA(ctx,cudaEventElapsedTime(&ms,e0,e1));
  *ms_per_launch=ms/float(reps);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return B2S_OK;
}

// ---- slab-mode step pieces (see include/b2s.h; orchestration in dualsphysics_b200/slab.py) ------------------------------
int b2s_dtvariable_impl(b2s_ctx* ctx, int interstep);

extern "C" int b2s_slab_forces(b2s_ctx* ctx, int interstep){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  if(ctx->prm.tboundary==B2S_BC_MDBC && (ctx->prm.mdbc_corrector || interstep!=B2S_INTER_SYMCORRECTOR)){ int r=b2s_mdbc_impl(ctx); if(r)return r; }
  return b2s_forces_impl(ctx,-1);     // -1: no DtVariable yet (the maxima are all-reduced first)
}
extern "C" int b2s_slab_apply_dt(b2s_ctx* ctx, int interstep){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  return b2s_dtvariable_impl(ctx,interstep);
}
extern "C" int b2s_slab_integrate(b2s_ctx* ctx, int mode){
  if(!ctx || mode<0 || mode>2)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  if(mode==1)b2s_set_dt_device(ctx,0);        // dt=SymplecticDtPre at the start of the step
  else if(mode==0)b2s_set_dt_device(ctx,1);   // Verlet: dt=DtVariable(true)
  return b2s_integrate_impl(ctx,mode,0,true);
}
// tail of the step without the synchronisation (slab_comm.cu synchronises once, after the last exchange of the step)
int b2s_slab_step_end_enqueue(b2s_ctx* ctx){
  { int r=b2s_damping_impl(ctx); if(r)return r; }
  if(!ctx->motions.empty()){ int r=b2s_motion_impl(ctx,0,0); if(r)return r; }
  k_step_end<<<1,1,0,ctx->stream>>>(ctx->step,ctx->prm.tstep==B2S_STEP_SYMPLECTIC? 1: 0); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
extern "C" int b2s_slab_step_end(b2s_ctx* ctx, double* dt_out){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  { int r=b2s_slab_step_end_enqueue(ctx); if(r)return r; }
  int r=b2s_sync_info(ctx);
  if(dt_out)*dt_out=ctx->step_h->dt;
  return r;
}
