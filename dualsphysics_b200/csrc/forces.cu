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
#include "b2s_internal.h"
#include <float.h>

#define F_BS 128

struct ForcesArgs {
  const DivState* ds; StepState* st;
  const uint32_t* beg;            // begincell (CPU form)
  const uint32_t* dcell; const typecode* code;
  const float4* posp; const float4* velrho;
  const float* tau; float* tau_new;   // 6 floats/particle (xx,xy,xz,yy,yz,zz) or NULL
  const float* ftomassp;
  float4* acear;
};

__device__ __forceinline__ float b2s_rcp(float x){ float r; asm("rcp.approx.ftz.f32 %0, %1;":"=f"(r):"f"(x)); return r; }
__device__ __forceinline__ float b2s_rsqrt(float x){ float r; asm("rsqrt.approx.ftz.f32 %0, %1;":"=f"(r):"f"(x)); return r; }

// fac = |grad W|/r  (FunSphKernel.h:53-65 cubic, :218-223 Wendland)
template<int TKER> __device__ __forceinline__ float kernel_fac(float rr2, const DevCte& c){
  const float rad=rr2*b2s_rsqrt(rr2);
  const float qq=rad*c.ovkernelh;
  if(TKER==B2S_KERNEL_WENDLAND){
    const float wqq1=1.f-0.5f*qq;
    return c.bwenh*wqq1*wqq1*wqq1;
  }
  else{
    if(rad>c.kernelh){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; return c.cub_c2*wqq2*b2s_rcp(rad); }
    else{ const float wqq2=qq*qq; return (c.cub_c1*qq+c.cub_d1*wqq2)*b2s_rcp(rad); }
  }
}
// Cubic-spline tensile correction (FunSphKernel.h:88-97)
__device__ __forceinline__ float cubic_tensil(float rr2,float rhop1,float pressp1,float rhop2,float pressp2,const DevCte& c){
  const float rad=rr2*b2s_rsqrt(rr2);
  const float qq=rad*c.ovkernelh;
  float wab;
  if(rad>c.kernelh){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; wab=c.cub_a24*(wqq2*wqq1); }
  else{ const float wqq2=qq*qq; wab=c.cub_a2*(1.0f+(0.75f*qq-1.5f)*wqq2); }
  float fab=wab*c.cub_od_wdeltap;
  fab*=fab; fab*=fab;
  const float tensilp1=(pressp1/(rhop1*rhop1))*(pressp1>0? 0.01f: -0.2f);
  const float tensilp2=(pressp2/(rhop2*rhop2))*(pressp2>0? 0.01f: -0.2f);
  return fab*(tensilp1+tensilp2);
}

struct PairAcc {
  float ax,ay,az,ar,visc,delta;
  float sxx,sxy,sxz,syy,syz,szz;   // 2*strain (SPS)
};

// One fluid/floating particle p1 against one neighbour p2 (JSphCpu.cpp:822-1041).
template<int TKER,int TVISCO,int TDENS,bool FLOATING,bool BOUNDP2>
__device__ __forceinline__ void pair_fluid(const DevCte& c, float drx,float dry,float drz,float rr2,
  const float4 velrhop1,float pressp1,bool ftp1,const float* taup1,
  const float4 velrhop2,float pressp2,typecode codep2,const float* __restrict__ taug,uint32_t p2,
  const float* __restrict__ ftomassp,float visco,PairAcc& a)
{
  const float fac=kernel_fac<TKER>(rr2,c);
  const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
  float massp2=(BOUNDP2? c.massb: c.massf);
  bool ftp2=false;
  if(FLOATING){
    ftp2=B2S_CODE_ISFLOATING(codep2);
    if(ftp2)massp2=ftomassp[B2S_CODE_TYPEVALUE(codep2)];
    if(ftp2 && TDENS==B2S_DDT_MOLTENI && massp2<=(c.massf*1.2f))a.delta=FLT_MAX;   // DELTA_HEAVYFLOATING (DualSphDef.h:104)
  }
  const float rhop1=velrhop1.w, rhop2=velrhop2.w;
  const float ovrhop2=b2s_rcp(rhop2);
  //-Momentum equation.
  {
    float prs=(pressp1+pressp2)*(b2s_rcp(rhop1)*ovrhop2);
    if(TKER==B2S_KERNEL_CUBIC)prs+=cubic_tensil(rr2,rhop1,pressp1,rhop2,pressp2,c);
    const float p_vpm=-prs*massp2;
    a.ax+=p_vpm*frx; a.ay+=p_vpm*fry; a.az+=p_vpm*frz;
  }
  //-Continuity equation.
  const float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
  a.ar+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhop1*ovrhop2);
  const float ovrr2eta=b2s_rcp(rr2+c.eta2);
  const float dot3=(TDENS!=B2S_DDT_NONE || TVISCO!=B2S_VISCO_ARTIFICIAL? drx*frx+dry*fry+drz*frz: 0.f);
  //-Density diffusion: Molteni & Colagrossi (JSphCpu.cpp:934-941).
  if(TDENS==B2S_DDT_MOLTENI && a.delta!=FLT_MAX){
    const float rhop1over2=rhop1*ovrhop2;
    const float visc_densi=c.ddtkh*c.cs0*(rhop1over2-1.f)*ovrr2eta;
    const float delta=visc_densi*dot3*massp2;
    a.delta=(BOUNDP2 && c.tboundary==B2S_BC_DBC? FLT_MAX: a.delta+delta);
  }
  //-Density diffusion: Fourtakas et al. (JSphCpu.cpp:943-950).
  if((TDENS==B2S_DDT_FOURTAKAS || (TDENS==B2S_DDT_FOURTAKASFULL && !BOUNDP2)) && a.delta!=FLT_MAX && !ftp2){
    const float rh=1.f+c.ddtgz*drz;
    // rh^(1/gamma) as fl32(1+series(rh-1)): |rh-1|<~1e-2, series error < 1e-9 relative, then the same
    // fl32 rounding and rho0*pw-rho0 cancellation as the reference's pow() ("DDT2 hydrostatic term").
    const float x=rh-1.f;
    const float pw=1.f+x*(c.ddt_c1+x*(c.ddt_c2+x*c.ddt_c3));
    const float drho=c.rhopzero*pw-c.rhopzero;
    const float visc_densi=c.ddtkh*c.cs0*((rhop2-rhop1)-drho)*ovrr2eta;
    const float delta=visc_densi*dot3*massp2*ovrhop2;
    a.delta=(BOUNDP2? FLT_MAX: a.delta-delta);
  }
  //-Viscosity.
  {
    const float dot=drx*dvx+dry*dvy+drz*dvz;
    const float dot_rr2=dot*ovrr2eta;
    a.visc=fmaxf(dot_rr2,a.visc);
    if(TVISCO==B2S_VISCO_ARTIFICIAL){
      if(dot<0){
        const float amubar=c.kernelh*dot_rr2;
        const float robar=(rhop1+rhop2)*0.5f;
        const float pi_visc=(-visco*c.cs0*amubar*b2s_rcp(robar))*massp2;
        a.ax-=pi_visc*frx; a.ay-=pi_visc*fry; a.az-=pi_visc*frz;
      }
    }
    else{
      const float robar2=(rhop1+rhop2);
      const float temp=4.f*visco*b2s_rcp((rr2+c.eta2)*robar2);
      const float vtemp=massp2*temp*dot3;
      a.ax+=vtemp*dvx; a.ay+=vtemp*dvy; a.az+=vtemp*dvz;
    }
    if(TVISCO==B2S_VISCO_LAMINARSPS){
      float sxx=taup1[0],sxy=taup1[1],sxz=taup1[2],syy=taup1[3],syz=taup1[4],szz=taup1[5];
      if(!BOUNDP2 && !ftp2){
        const float2* t2=(const float2*)(taug+size_t(p2)*6);
        const float2 t0=t2[0],t1=t2[1],t3=t2[2];
        sxx+=t0.x; sxy+=t0.y; sxz+=t1.x; syy+=t1.y; syz+=t3.x; szz+=t3.y;
      }
      a.ax+=massp2*(sxx*frx+sxy*fry+sxz*frz);
      a.ay+=massp2*(sxy*frx+syy*fry+syz*frz);
      a.az+=massp2*(sxz*frx+syz*fry+szz*frz);
      if(!ftp1){
        const float volp2=-massp2*ovrhop2;
        float dv=dvx*volp2; a.sxx+=dv*frx; a.sxy+=dv*fry; a.sxz+=dv*frz;
              dv=dvy*volp2; a.sxy+=dv*frx; a.syy+=dv*fry; a.syz+=dv*frz;
              dv=dvz*volp2; a.sxz+=dv*frx; a.syz+=dv*fry; a.szz+=dv*frz;
      }
    }
  }
}

// Neighbour cell ranges of cunsearch::InitCte (JCellSearch_iker.h:34-57), cells relative to cellmin.
struct CellRange { int cx,cy,cz,xini,xfin,yini,yfin,zini,zfin; };
__device__ __forceinline__ CellRange cell_range(uint32_t dcell,const DevCte& c,const DivState* ds){
  CellRange r;
  r.cx=int(B2S_DCEL_CELLX(c.domcellcode,dcell))-int(ds->cellmin[0]);
  r.cy=int(B2S_DCEL_CELLY(c.domcellcode,dcell))-int(ds->cellmin[1]);
  r.cz=int(B2S_DCEL_CELLZ(c.domcellcode,dcell))-int(ds->cellmin[2]);
  const int d=c.scelldiv, ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz);
  r.xini=r.cx-min(r.cx,d); r.xfin=r.cx+min(ncx-r.cx-1,d)+1;
  r.yini=r.cy-min(r.cy,d); r.yfin=r.cy+min(ncy-r.cy-1,d)+1;
  r.zini=r.cz-min(r.cz,d); r.zfin=r.cz+min(ncz-r.cz-1,d)+1;
  return r;
}

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
  if(p1<np){
    const float4 pp1=g.posp[p1];
    const float4 velrhop1=g.velrho[p1];
    const float pressp1=pp1.w;
    bool ftp1=false;
    if(FLOATING)ftp1=B2S_CODE_ISFLOATING(g.code[p1]);
    float taup1[6]={0,0,0,0,0,0};
    if(TVISCO==B2S_VISCO_LAMINARSPS){
      const float2* t=(const float2*)(g.tau+size_t(p1)*6);
      const float2 t0=t[0],t1=t[1],t2=t[2];
      taup1[0]=t0.x; taup1[1]=t0.y; taup1[2]=t1.x; taup1[3]=t1.y; taup1[4]=t2.x; taup1[5]=t2.y;
    }
    const CellRange r=cell_range(g.dcell[p1],c,ds);
    const int ncx=int(ds->ncx), nsheet=int(ds->nsheet);
    PairAcc af={0,0,0,0,0,0,0,0,0,0,0,0};
    PairAcc ab={0,0,0,0,0,0,0,0,0,0,0,0};
    if(FLOATING && ftp1 && TDENS!=B2S_DDT_NONE){ af.delta=FLT_MAX; ab.delta=FLT_MAX; }
    //-Fluid/floating neighbours.
    {
      const uint32_t* beg=g.beg+ds->box_fluid;
      for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
        const int row=y*ncx+z*nsheet;
        const float offy=c.scell*float(r.cy-y), offz=c.scell*float(r.cz-z);
        for(int x=r.xini;x<r.xfin;x++){
          const uint32_t pini=beg[row+x], pfin=beg[row+x+1];
          const float offx=c.scell*float(r.cx-x);
          for(uint32_t p2=pini;p2<pfin;p2++){
            const float4 pp2=g.posp[p2];
            const float drx=(pp1.x-pp2.x)+offx, dry=(pp1.y-pp2.y)+offy, drz=(pp1.z-pp2.z)+offz;
            const float rr2=drx*drx+dry*dry+drz*drz;
            if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO){
              const typecode cod2=(FLOATING? g.code[p2]: typecode(0));
              pair_fluid<TKER,TVISCO,TDENS,FLOATING,false>(c,drx,dry,drz,rr2,velrhop1,pressp1,ftp1,taup1,
                g.velrho[p2],pp2.w,cod2,g.tau,p2,g.ftomassp,c.visco_f,af);
            }
          }
        }
      }
    }
    //-Boundary neighbours.
    {
      const uint32_t* beg=g.beg;
      for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
        const int row=y*ncx+z*nsheet;
        const float offy=c.scell*float(r.cy-y), offz=c.scell*float(r.cz-z);
        for(int x=r.xini;x<r.xfin;x++){
          const uint32_t pini=beg[row+x], pfin=beg[row+x+1];
          const float offx=c.scell*float(r.cx-x);
          for(uint32_t p2=pini;p2<pfin;p2++){
            const float4 pp2=g.posp[p2];
            const float drx=(pp1.x-pp2.x)+offx, dry=(pp1.y-pp2.y)+offy, drz=(pp1.z-pp2.z)+offz;
            const float rr2=drx*drx+dry*dry+drz*drz;
            if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO){
              pair_fluid<TKER,TVISCO,TDENS,FLOATING,true>(c,drx,dry,drz,rr2,velrhop1,pressp1,ftp1,taup1,
                g.velrho[p2],pp2.w,typecode(0),g.tau,p2,g.ftomassp,c.visco_b,ab);
            }
          }
        }
      }
    }
    //-Store (JSphCpu.cpp:1086-1102 fluid pass then bound pass; JSphGpu_ker.cu:939-950 for the GPU semantics).
    // The density-diffusion sum is kept apart and added last, and a DBC boundary neighbour (FLT_MAX marker)
    // discards it completely: this is what both reference builds do (GPU: one running deltap1, JSphGpu_ker.cu:939-946;
    // CPU: DDTArray=true on CPU, JSph.cpp:927, Delta_c added after both passes, JSphCpuSingle.cpp:723-731).
    float4 res;
    res.x=af.ax+ab.ax; res.y=af.ay+ab.ay; res.z=af.az+ab.az; res.w=af.ar+ab.ar;
    if(TDENS!=B2S_DDT_NONE && af.delta!=FLT_MAX && ab.delta!=FLT_MAX)res.w+=(af.delta+ab.delta);
    if(c.simulate2d)res.y=0.f;
    g.acear[p1]=res;
    viscmax=fmaxf(af.visc,ab.visc);
    if(B2S_CODE_ISNORMAL(g.code[p1]))ace2=res.x*res.x+res.y*res.y+res.z*res.z;
    if(TVISCO==B2S_VISCO_LAMINARSPS){
      // ComputeSpsTau (JSphCpu.cpp:1279-1307) on the accumulated 2*strain of both passes.
      const float sxx=af.sxx+ab.sxx, sxy=af.sxy+ab.sxy, sxz=af.sxz+ab.sxz, syy=af.syy+ab.syy, syz=af.syz+ab.syz, szz=af.szz+ab.szz;
      const float pow1=sxx*sxx+syy*syy+szz*szz;
      const float prr=sxy*sxy+sxz*sxz+syz*syz+pow1+pow1;
      const float visc_sps=c.spssmag*sqrtf(prr);
      const float div_u=sxx+syy+szz;
      const float sps_k=visc_sps*div_u;
      const float sps_blin=c.spsblin*prr;
      const float sumsps=-(sps_k+sps_blin);
      const float twovisc_sps=visc_sps+visc_sps;
      const float one_rho=1.0f/velrhop1.w;
      float2* tn=(float2*)(g.tau_new+size_t(p1)*6);
      tn[0]=make_float2(one_rho*(twovisc_sps*sxx+sumsps),one_rho*(visc_sps*sxy));
      tn[1]=make_float2(one_rho*(visc_sps*sxz),one_rho*(twovisc_sps*syy+sumsps));
      tn[2]=make_float2(one_rho*(visc_sps*syz),one_rho*(twovisc_sps*szz+sumsps));
    }
  }
  //-Fused maxima: ViscDtMax and AceMax^2 (replaces ReduMaxFloat x2 + cudaMemcpy, JSphGpu_ker.cu:75-125).
  #pragma unroll
  for(int o=16;o;o>>=1){ viscmax=fmaxf(viscmax,__shfl_xor_sync(0xffffffffu,viscmax,o)); ace2=fmaxf(ace2,__shfl_xor_sync(0xffffffffu,ace2,o)); }
  if((threadIdx.x&31)==0){
    if(viscmax>0.f)atomicMax(&g.st->viscdtmax_bits,__float_as_uint(viscmax));
    if(ace2>0.f)atomicMax(&g.st->acemax2_bits,__float_as_uint(ace2));
  }
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
      const int ncx=int(ds->ncx), nsheet=int(ds->nsheet);
      const uint32_t* beg=g.beg+ds->box_fluid;
      for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
        const int row=y*ncx+z*nsheet;
        const float offy=c.scell*float(r.cy-y), offz=c.scell*float(r.cz-z);
        for(int x=r.xini;x<r.xfin;x++){
          const uint32_t pini=beg[row+x], pfin=beg[row+x+1];
          const float offx=c.scell*float(r.cx-x);
          for(uint32_t p2=pini;p2<pfin;p2++){
            const float4 pp2=g.posp[p2];
            const float drx=(pp1.x-pp2.x)+offx, dry=(pp1.y-pp2.y)+offy, drz=(pp1.z-pp2.z)+offz;
            const float rr2=drx*drx+dry*dry+drz*drz;
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
    }
    g.acear[p1]=make_float4(0.f,0.f,0.f,arp1);
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

int b2s_forces_tile_launch(b2s_ctx* ctx, const ForcesArgs& g);   // forces_tile.cu

int b2s_forces_fluid_only(b2s_ctx* ctx){
  ForcesArgs g;
  g.ds=ctx->div; g.st=ctx->step; g.beg=(const uint32_t*)ctx->begincell; g.dcell=ctx->cur.dcell; g.code=ctx->cur.code;
  g.posp=ctx->posp; g.velrho=ctx->cur.velrho; g.tau=ctx->spstau; g.tau_new=ctx->spstau_new; g.ftomassp=ctx->ft_massp_dev; g.acear=ctx->acear;
  const uint32_t npf=(ctx->np_host>ctx->npb_host? ctx->np_host-ctx->npb_host: 0u);
  if(!npf)return B2S_OK;
  if(ctx->variant==1)return b2s_forces_tile_launch(ctx,g);
  fluid_kernel_t k=b2s_pick_fluid_gather(ctx->cte);
  k<<<b2s_blocks(npf,F_BS),F_BS,0,ctx->stream>>>(g,ctx->cte); B2S_LAUNCHED(ctx);
  return B2S_OK;
}

int b2s_forces_impl(b2s_ctx* ctx, int interstep){
  cudaStream_t s=ctx->stream;
  ForcesArgs g;
  g.ds=ctx->div; g.st=ctx->step; g.beg=(const uint32_t*)ctx->begincell; g.dcell=ctx->cur.dcell; g.code=ctx->cur.code;
  g.posp=ctx->posp; g.velrho=ctx->cur.velrho; g.tau=ctx->spstau; g.tau_new=ctx->spstau_new; g.ftomassp=ctx->ft_massp_dev; g.acear=ctx->acear;
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
  k_dtvariable<<<1,1,0,s>>>(ctx->step,(interstep==B2S_INTER_SYMPREDICTOR? 0: 1),ctx->cte); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
