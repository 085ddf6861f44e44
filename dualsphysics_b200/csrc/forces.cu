// forces.cu -- particle interaction (fluid<-fluid/bound, bound<-fluid), SPS tau and the fused dt maxima.
//
// Replaces (all under /root/reference/src/source):
//   JSphGpu::PreInteraction_Forces            JSphGpu.cpp:846-913   (memsets + VelMax; folded away / into k_reorder)
//   cusph::Interaction_Forces dispatch         JSphGpu_ker.cu:1043-1179
//   KerInteractionForcesFluid(+Box)            JSphGpu_ker.cu:539-990
//   KerInteractionForcesBound(+Box)            JSphGpu_ker.cu:450-533
//   cusph::ComputeSpsTau                       JSphGpu_ker.cu:1387-1436
//   ComputeAceMod / ReduMaxFloat / ComputeAceMax  JSphGpu_ker.cu:75-312, JSphGpuSingle.cpp:719-727
//   JSphGpu::DtVariable                        JSphGpu.cpp:1103-1147  (k_dtvariable, device side)
// Arithmetic follows the reference's CPU statement of the same loop (JSphCpu.cpp:640-1125), which is the
// runnable oracle: per-particle pressure (Press_c, JSphCpu.cpp:537-547) instead of a pow() per pair,
// fluid pass then boundary pass accumulated separately and added in that order.
//
// Two fluid kernels:
//   variant 0 "gather": one thread per particle walking the 3x3(x5x5) cell rows in the reference order
//             (bit-comparable summation order; the parity anchor).
//   variant 1 "tile":   see forces_tile.cu (shared-memory staged, sm_100a tuned).
#include "forces_common.cuh"

// ------------------------------------------------------------------------------------------------
// Variant 0: gather kernel, one thread per fluid/floating particle.
// ------------------------------------------------------------------------------------------------
template<int TKER,int TVISCO,int TDENS,bool FLOATING>
__global__ void __launch_bounds__(F_BS) k_forces_fluid_gather(ForcesArgs g, const __grid_constant__ DevCte c)
{
  const DivState* ds=g.ds;
  const uint32_t npb=ds->npb, np=ds->np;
  const uint32_t p1=npb+blockIdx.x*blockDim.x+threadIdx.x;
  float ace2=0.f, viscmax=0.f;
  const bool valid1=(p1<np && B2S_CODE_ISNORMAL(g.code[p1]));   // slab halo ghosts (CODE_PERIODIC) are neighbours only
  // boundary pass only for warps with a boundary particle inside some neighbour window (per-cell flags of the divide)
  bool nearb=(g.nearbound==nullptr);
  if(valid1 && !nearb){ const CellRange q=cell_range(g.dcell[p1],c,ds); nearb=(g.nearbound[q.cx+q.cy*int(ds->ncx)+q.cz*int(ds->nsheet)]!=0); }
  const bool warpnearb=(__any_sync(0xffffffffu,nearb)!=0);
  if(valid1)gather_particle<TKER,TVISCO,TDENS,FLOATING>(g,c,ds,p1,warpnearb,ace2,viscmax);
  //-Fused maxima: ViscDtMax and AceMax^2 (replaces ReduMaxFloat x2 + cudaMemcpy, JSphGpu_ker.cu:75-125).
  reduce_maxima(g.st,viscmax,ace2);
}

// ------------------------------------------------------------------------------------------------
// Boundary <- fluid (KerInteractionForcesBound, JSphGpu_ker.cu:450-533; CPU JSphCpu.cpp:640-700).
// Runs over [0,npb): particles beyond npbok (outside the fluid cell domain) get ar=0.
// ------------------------------------------------------------------------------------------------
template<int TKER,bool FLOATING>
__global__ void __launch_bounds__(F_BS) k_forces_bound(ForcesArgs g, const __grid_constant__ DevCte c)
{
  const DivState* ds=g.ds;
  const uint32_t p1=blockIdx.x*blockDim.x+threadIdx.x;
  float visc=0.f;
  if(p1<ds->npb){
    float arp1=0.f;
    if(p1<ds->npbok){
      const float4 pp1=g.posp[p1];
      const float4 velrhop1=g.velrho[p1];
      const CellRange r=cell_range(g.dcell[p1],c,ds);
      const int ncx=int(ds->ncx), nsheet=int(ds->nsheet), cmx=int(ds->cellmin[0]);
      const uint32_t* beg=g.beg+ds->box_fluid;
      // A row of neighbour cells is ONE index range (x-contiguous cells, block-relative x: at most two x offsets, b2s_internal.h
      // B2S_XBLK), settled with two loads when it holds no fluid (most of a tank's walls) and trimmed to the cells the kernel sphere
      // reaches.  Same particles in the same order as the cell-by-cell walk: the trim only drops cells beyond 2h.
      const float xin=pp1.x-c.scell*float((r.cx+cmx)&((1<<B2S_XBLK)-1));
      const bool oneblk=(((r.xini+cmx)>>B2S_XBLK)==((r.xfin-1+cmx)>>B2S_XBLK));
      const int xmid0=(((r.xfin-1+cmx)>>B2S_XBLK)<<B2S_XBLK)-cmx;          // first cell of the second block
      const float offxA=b2s_offx(c,r.cx+cmx,r.xini+cmx), offxB=b2s_offx(c,r.cx+cmx,r.xfin-1+cmx);
      for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
        const uint32_t* brow=beg+(y*ncx+z*nsheet);
        if(brow[r.xini]==brow[r.xfin])continue;
        const float offy=c.scell*float(r.cy-y), offz=c.scell*float(r.cz-z);
        int xa,xb;
        if(!row_trim(c,r,xin,FADD(pp1.y,offy),FADD(pp1.z,offz),xa,xb))continue;
        const uint32_t pini=brow[xa], pfin=brow[xb];
        const uint32_t pmid=(oneblk || xmid0<=xa? pini: (xmid0>=xb? pfin: brow[xmid0]));
        for(uint32_t p2=pini;p2<pfin;p2++){
          const float4 pp2=g.posp[p2];
          const float drx=b2s_dr(pp1.x,pp2.x,(p2>=pmid? offxB: offxA)), dry=b2s_dr(pp1.y,pp2.y,offy), drz=b2s_dr(pp1.z,pp2.z,offz);
          const float rr2=b2s_rr2(drx,dry,drz);
          if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO){
            const float fac=kernel_fac<TKER>(rr2,c);
            const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
            float massp2=c.massf;
            if(FLOATING){ const typecode cod=g.code[p2]; if(B2S_CODE_ISFLOATING(cod))massp2=g.ftomassp[B2S_CODE_TYPEVALUE(cod)]; }
            const float4 velrhop2=g.velrho[p2];
            const float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
            arp1+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(velrhop1.w*b2s_rcp(velrhop2.w));
            const float dot=drx*dvx+dry*dvy+drz*dvz;
            visc=fmaxf(dot*b2s_rcp(rr2+c.eta2),visc);
          }
        }
      }
    }
    g.acear[p1]=make_float4(0.f,0.f,0.f,arp1);
    if(!B2S_CODE_ISNORMAL(g.code[p1]))visc=0.f;
  }
  #pragma unroll
  for(int o=16;o;o>>=1)visc=fmaxf(visc,__shfl_xor_sync(0xffffffffu,visc,o));
  if((threadIdx.x&31)==0 && visc>0.f)atomicMax(&g.st->viscdtmax_bits,__float_as_uint(visc));
}

// ------------------------------------------------------------------------------------------------
// DtVariable on the device (JSphGpu.cpp:1103-1147 / JSphCpu.cpp:2031-2076), one thread.
// ------------------------------------------------------------------------------------------------
__global__ void k_dtvariable(StepState* st, int final_, const __grid_constant__ DevCte c){
  if(threadIdx.x||blockIdx.x)return;
  const float velmax=sqrtf(__uint_as_float(st->velmax2_bits));           // CalcVelMaxSeq: sqrt(float)
  const double acemax=sqrt(double(__uint_as_float(st->acemax2_bits)));    // ComputeAceMaxSeq: sqrt(double(float))
  const float viscdtmax=__uint_as_float(st->viscdtmax_bits);
  st->velmax=velmax; st->acemax=acemax; st->viscdtmax=viscdtmax;
  const double dt1=(acemax? sqrt(c.kernelh_d/acemax): DBL_MAX);
  const double dt2=c.kernelh_d/(fmax(c.cs0_d,double(velmax)*10.)+c.kernelh_d*double(viscdtmax));
  double dt=c.cfl*fmin(dt1,dt2);
  if(isnan(dt) || isinf(dt))st->error=B2S_ERR_DTINVALID;
  if(dt<c.dtmin){ dt=c.dtmin; st->dtmodif++; }
  if(final_==0)st->dt_p=dt; else st->dt_c=dt;
  st->dt_last=dt;
}

// ------------------------------------------------------------------------------------------------
typedef void (*fluid_kernel_t)(ForcesArgs, const DevCte);

template<int TKER,int TVISCO,int TDENS> static fluid_kernel_t pick_ft(bool ft){
  return ft? (fluid_kernel_t)k_forces_fluid_gather<TKER,TVISCO,TDENS,true>: (fluid_kernel_t)k_forces_fluid_gather<TKER,TVISCO,TDENS,false>;
}
template<int TKER,int TVISCO> static fluid_kernel_t pick_dens(int td,bool ft){
  switch(td){
    case B2S_DDT_NONE: return pick_ft<TKER,TVISCO,B2S_DDT_NONE>(ft);
    case B2S_DDT_MOLTENI: return pick_ft<TKER,TVISCO,B2S_DDT_MOLTENI>(ft);
    case B2S_DDT_FOURTAKAS: return pick_ft<TKER,TVISCO,B2S_DDT_FOURTAKAS>(ft);
    default: return pick_ft<TKER,TVISCO,B2S_DDT_FOURTAKASFULL>(ft);
  }
}
template<int TKER> static fluid_kernel_t pick_visco(int tv,int td,bool ft){
  switch(tv){
    case B2S_VISCO_ARTIFICIAL: return pick_dens<TKER,B2S_VISCO_ARTIFICIAL>(td,ft);
    case B2S_VISCO_LAMINARSPS: return pick_dens<TKER,B2S_VISCO_LAMINARSPS>(td,ft);
    default: return pick_dens<TKER,B2S_VISCO_LAMINAR>(td,ft);
  }
}
fluid_kernel_t b2s_pick_fluid_gather(const DevCte& c){
  const bool ft=(c.use_floating!=0);
  return c.tkernel==B2S_KERNEL_WENDLAND? pick_visco<B2S_KERNEL_WENDLAND>(c.tvisco,c.tdensity,ft): pick_visco<B2S_KERNEL_CUBIC>(c.tvisco,c.tdensity,ft);
}

int b2s_forces_pair2_launch(b2s_ctx* ctx, const ForcesArgs& g);  // forces_pair2.cu
bool b2s_forces_pair2_supported(const DevCte& c);
int b2s_forces_tile_launch(b2s_ctx* ctx, const ForcesArgs& g);   // forces_tile.cu
bool b2s_forces_tile_supported(const DevCte& c);

int b2s_forces_fluid_only(b2s_ctx* ctx){
  ForcesArgs g;
  g.ds=ctx->div; g.st=ctx->step; g.beg=(const uint32_t*)ctx->begincell; g.dcell=ctx->cur.dcell; g.code=ctx->cur.code;
  g.posp=ctx->posp; g.posh=ctx->posh; g.boundmode=ctx->boundmode; g.tangenvel=ctx->tangenvel; g.nearbound=(ctx->nearbound_valid? ctx->nearbound: NULL); g.velrho=ctx->cur.velrho; g.tau=ctx->spstau; g.tau_new=ctx->spstau_new; g.ftomassp=ctx->ft_massp_dev; g.acear=ctx->acear;
  const uint32_t npf=ctx->npf_host;
  if(!npf)return B2S_OK;
  if(ctx->variant==6 && b2s_forces_tile_supported(ctx->cte))return b2s_forces_tile_launch(ctx,g);
  if((ctx->variant==5 || ctx->variant==6) && b2s_forces_pair2_supported(ctx->cte))return b2s_forces_pair2_launch(ctx,g);
  fluid_kernel_t k=b2s_pick_fluid_gather(ctx->cte);
  k<<<b2s_blocks(npf,F_BS),F_BS,0,ctx->stream>>>(g,ctx->cte); B2S_LAUNCHED(ctx);
  return B2S_OK;
}

int b2s_forces_impl(b2s_ctx* ctx, int interstep){
  cudaStream_t s=ctx->stream;
  ForcesArgs g;
  g.ds=ctx->div; g.st=ctx->step; g.beg=(const uint32_t*)ctx->begincell; g.dcell=ctx->cur.dcell; g.code=ctx->cur.code;
  g.posp=ctx->posp; g.posh=ctx->posh; g.boundmode=ctx->boundmode; g.tangenvel=ctx->tangenvel; g.nearbound=(ctx->nearbound_valid? ctx->nearbound: NULL); g.velrho=ctx->cur.velrho; g.tau=ctx->spstau; g.tau_new=ctx->spstau_new; g.ftomassp=ctx->ft_massp_dev; g.acear=ctx->acear;
  int r=b2s_forces_fluid_only(ctx); if(r)return r;
  if(ctx->npb_host){
    const bool ft=(ctx->cte.use_floating!=0);
    const unsigned gb=b2s_blocks(ctx->npb_host,F_BS);
    if(ctx->cte.tkernel==B2S_KERNEL_WENDLAND){
      if(ft)k_forces_bound<B2S_KERNEL_WENDLAND,true><<<gb,F_BS,0,s>>>(g,ctx->cte); else k_forces_bound<B2S_KERNEL_WENDLAND,false><<<gb,F_BS,0,s>>>(g,ctx->cte);
    }
    else{
      if(ft)k_forces_bound<B2S_KERNEL_CUBIC,true><<<gb,F_BS,0,s>>>(g,ctx->cte); else k_forces_bound<B2S_KERNEL_CUBIC,false><<<gb,F_BS,0,s>>>(g,ctx->cte);
    }
    B2S_LAUNCHED(ctx);
  }
  if(ctx->cte.tvisco==B2S_VISCO_LAMINARSPS)std::swap(ctx->spstau,ctx->spstau_new);
  if(interstep>=0){ k_dtvariable<<<1,1,0,s>>>(ctx->step,(interstep==B2S_INTER_SYMPREDICTOR? 0: 1),ctx->cte); B2S_LAUNCHED(ctx); }
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}

// DtVariable alone (slab mode: called after the maxima have been all-reduced over the ranks).
int b2s_dtvariable_impl(b2s_ctx* ctx, int interstep){
  k_dtvariable<<<1,1,0,ctx->stream>>>(ctx->step,(interstep==B2S_INTER_SYMPREDICTOR? 0: 1),ctx->cte); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
