// forces_common.cuh -- device code shared by the two fluid interaction kernels (forces.cu gather, forces_tile.cu tile):
// kernel gradients, the per-pair body and the neighbour cell ranges.  Arithmetic follows the reference's CPU statement
// (src/source/JSphCpu.cpp:736-1125, FunSphKernel.h, FunSphEos.h); see forces.cu for the full citation list.
#pragma once
#include "b2s_internal.h"
#include <float.h>

#define F_BS 128

struct ForcesArgs {
  const DivState* ds; StepState* st;
  const uint32_t* beg;            // begincell (CPU form)
  const uint32_t* dcell; const typecode* code;
  const float4* posp; const float4* velrho; const uint4* posh;
  const float* tau; float* tau_new;   // 6 floats/particle (xx,xy,xz,yy,yz,zz) or NULL
  const float* ftomassp;
  float4* acear;
  const uint8_t* boundmode; const float4* tangenvel;   // mDBC2 (SlipMode 2/3), else NULL
  const uint8_t* nearbound;       // per cell: boundary within the neighbour window (celldiv.cu k_nearbound_mark), or NULL
};

__device__ __forceinline__ float b2s_rcp(float x){ float r; asm("rcp.approx.ftz.f32 %0, %1;":"=f"(r):"f"(x)); return r; }
__device__ __forceinline__ float b2s_rsqrt(float x){ float r; asm("rsqrt.approx.ftz.f32 %0, %1;":"=f"(r):"f"(x)); return r; }

// fac = |grad W|/r  (FunSphKernel.h:53-65 cubic, :218-223 Wendland)
template<int TKER> __device__ __forceinline__ float kernel_fac(float rr2, const DevCte& c){
  const float rad=__fmul_rn(rr2,b2s_rsqrt(rr2));
  const float qq=__fmul_rn(rad,c.ovkernelh);
  if(TKER==B2S_KERNEL_WENDLAND){
    const float wqq1=__fmaf_rn(-0.5f,qq,1.f);
    return __fmul_rn(__fmul_rn(__fmul_rn(c.bwenh,wqq1),wqq1),wqq1);
  }
  else{
    if(rad>c.kernelh){ const float wqq1=__fsub_rn(2.0f,qq); const float wqq2=__fmul_rn(wqq1,wqq1); return __fmul_rn(__fmul_rn(c.cub_c2,wqq2),b2s_rcp(rad)); }
    else{ const float wqq2=__fmul_rn(qq,qq); return __fmul_rn(__fmaf_rn(c.cub_d1,wqq2,__fmul_rn(c.cub_c1,qq)),b2s_rcp(rad)); }
  }
}
// Cubic-spline tensile correction (FunSphKernel.h:88-97)
__device__ __forceinline__ float cubic_tensil(float rr2,float rhop1,float pressp1,float rhop2,float pressp2,const DevCte& c){
  const float rad=__fmul_rn(rr2,b2s_rsqrt(rr2));
  const float qq=__fmul_rn(rad,c.ovkernelh);
  float wab;
  if(rad>c.kernelh){ const float wqq1=__fsub_rn(2.0f,qq); const float wqq2=__fmul_rn(wqq1,wqq1); wab=__fmul_rn(c.cub_a24,__fmul_rn(wqq2,wqq1)); }
  else{ const float wqq2=__fmul_rn(qq,qq); wab=__fmul_rn(c.cub_a2,__fmaf_rn(__fmaf_rn(0.75f,qq,-1.5f),wqq2,1.0f)); }
  float fab=__fmul_rn(wab,c.cub_od_wdeltap);
  fab=__fmul_rn(fab,fab); fab=__fmul_rn(fab,fab);
  const float tensilp1=__fmul_rn(__fmul_rn(pressp1,b2s_rcp(__fmul_rn(rhop1,rhop1))),(pressp1>0? 0.01f: -0.2f));
  const float tensilp2=__fmul_rn(__fmul_rn(pressp2,b2s_rcp(__fmul_rn(rhop2,rhop2))),(pressp2>0? 0.01f: -0.2f));
  return __fmul_rn(fab,__fadd_rn(tensilp1,tensilp2));
}

struct PairAcc {
  float ax,ay,az,ar,visc,delta;
  float sxx,sxy,sxz,syy,syz,szz;   // 2*strain (SPS)
};

// Explicit-rounding helpers: the pair arithmetic is written with intrinsics so that every kernel variant executes
// exactly the same operation sequence (nvcc never re-associates or re-contracts __f*_rn), which is what makes the
// variants bit-identical to each other.
#define FMUL(a,b)   __fmul_rn((a),(b))
#define FADD(a,b)   __fadd_rn((a),(b))
#define FSUB(a,b)   __fsub_rn((a),(b))
#define FFMA(a,b,c) __fmaf_rn((a),(b),(c))
__device__ __forceinline__ float b2s_dot3(float ax,float ay,float az,float bx,float by,float bz){ return FFMA(az,bz,FFMA(ay,by,FMUL(ax,bx))); }
__device__ __forceinline__ float b2s_rr2(float dx,float dy,float dz){ return FFMA(dz,dz,FFMA(dy,dy,FMUL(dx,dx))); }
// dr=(a+off)-b: a+off only depends on the target and on the neighbour row, so every kernel hoists it out of its pair loop
__device__ __forceinline__ float b2s_dr(float a,float b,float off){ return FSUB(FADD(a,off),b); }

// One fluid/floating particle p1 against one neighbour p2 (JSphCpu.cpp:822-1041).
template<int TKER,int TVISCO,int TDENS,bool FLOATING,bool BOUNDP2>
__device__ __forceinline__ void pair_fluid(const DevCte& c, float drx,float dry,float drz,float rr2,
  const float4 velrhop1,float pressp1,bool ftp1,const float* taup1,
  const float4 velrhop2,float pressp2,typecode codep2,const float* __restrict__ taug,uint32_t p2,
  const float* __restrict__ ftomassp,float visco,PairAcc& a,const uint8_t* __restrict__ boundmode=nullptr,const float4* __restrict__ tangenvel=nullptr)
{
  const float fac=kernel_fac<TKER>(rr2,c);
  const float frx=FMUL(fac,drx), fry=FMUL(fac,dry), frz=FMUL(fac,drz);
  float massp2=(BOUNDP2? c.massb: c.massf);
  //-mDBC2: a boundary particle whose ghost node is not submerged has no mass (JSphCpu.cpp:851-853).
  const bool m2=(BOUNDP2 && c.mdbc2 && boundmode!=nullptr);
  if(m2 && boundmode[p2]==B2S_BMODE_MDBC2OFF)massp2=0.f;
  bool ftp2=false;
  if(FLOATING){
    ftp2=B2S_CODE_ISFLOATING(codep2);
    if(ftp2)massp2=ftomassp[B2S_CODE_TYPEVALUE(codep2)];
    if(ftp2 && TDENS==B2S_DDT_MOLTENI && massp2<=(c.massf*1.2f))a.delta=FLT_MAX;   // DELTA_HEAVYFLOATING (DualSphDef.h:104)
  }
  const float rhop1=velrhop1.w, rhop2=velrhop2.w;
  const float ovrhop2=b2s_rcp(rhop2);
  //-Momentum equation.
  float p_vpm;
  {
    float prs=FMUL(FADD(pressp1,pressp2),FMUL(b2s_rcp(rhop1),ovrhop2));
    if(TKER==B2S_KERNEL_CUBIC)prs=FADD(prs,cubic_tensil(rr2,rhop1,pressp1,rhop2,pressp2,c));
    p_vpm=-FMUL(prs,massp2);
    a.ax=FFMA(p_vpm,frx,a.ax); a.ay=FFMA(p_vpm,fry,a.ay); a.az=FFMA(p_vpm,frz,a.az);
  }
  //-Continuity equation.
  const float dvx=FSUB(velrhop1.x,velrhop2.x), dvy=FSUB(velrhop1.y,velrhop2.y), dvz=FSUB(velrhop1.z,velrhop2.z);
  a.ar=FFMA(FMUL(massp2,b2s_dot3(dvx,dvy,dvz,frx,fry,frz)),FMUL(rhop1,ovrhop2),a.ar);
  const float ovrr2eta=b2s_rcp(FADD(rr2,c.eta2));
  const float dot3=(TDENS!=B2S_DDT_NONE || TVISCO!=B2S_VISCO_ARTIFICIAL? b2s_dot3(drx,dry,drz,frx,fry,frz): 0.f);
  //-Density diffusion: Molteni & Colagrossi (JSphCpu.cpp:934-941).
  if(TDENS==B2S_DDT_MOLTENI && a.delta!=FLT_MAX){
    const float rhop1over2=FMUL(rhop1,ovrhop2);
    const float visc_densi=FMUL(FMUL(FMUL(c.ddtkh,c.cs0),FSUB(rhop1over2,1.f)),ovrr2eta);
    const float delta=FMUL(FMUL(visc_densi,dot3),massp2);
    a.delta=(BOUNDP2 && c.tboundary==B2S_BC_DBC? FLT_MAX: FADD(a.delta,delta));
  }
  //-Density diffusion: Fourtakas et al. (JSphCpu.cpp:943-950).
  if((TDENS==B2S_DDT_FOURTAKAS || (TDENS==B2S_DDT_FOURTAKASFULL && !BOUNDP2)) && a.delta!=FLT_MAX && !ftp2){
    const float rh=FFMA(c.ddtgz,drz,1.f);
    // rh^(1/gamma) as fl32(1+series(rh-1)): |rh-1|<~1e-2, series error < 1e-9 relative, then the same
    // fl32 rounding and rho0*pw-rho0 cancellation as the reference's pow() ("DDT2 hydrostatic term").
    const float x=FSUB(rh,1.f);
    const float pw=FFMA(x,FFMA(x,FFMA(x,c.ddt_c3,c.ddt_c2),c.ddt_c1),1.f);
    const float drho=FSUB(FMUL(c.rhopzero,pw),c.rhopzero);
    const float visc_densi=FMUL(FMUL(FMUL(c.ddtkh,c.cs0),FSUB(FSUB(rhop2,rhop1),drho)),ovrr2eta);
    const float delta=FMUL(FMUL(FMUL(visc_densi,dot3),massp2),ovrhop2);
    a.delta=(BOUNDP2? FLT_MAX: FSUB(a.delta,delta));
  }
  //-Viscosity.
  {
    //-mDBC2: the viscous terms see the tangential velocity of the boundary particle (JSphCpu.cpp:990-995).
    float dvx=FSUB(velrhop1.x,velrhop2.x), dvy=FSUB(velrhop1.y,velrhop2.y), dvz=FSUB(velrhop1.z,velrhop2.z);
    if(m2){ const float4 tv=tangenvel[p2]; dvx=FSUB(velrhop1.x,tv.x); dvy=FSUB(velrhop1.y,tv.y); dvz=FSUB(velrhop1.z,tv.z); }
    const float dot=b2s_dot3(drx,dry,drz,dvx,dvy,dvz);
    const float dot_rr2=FMUL(dot,ovrr2eta);
    a.visc=fmaxf(dot_rr2,a.visc);
    if(TVISCO==B2S_VISCO_ARTIFICIAL){
      if(dot<0){
        const float amubar=FMUL(c.kernelh,dot_rr2);
        const float robar=FMUL(FADD(rhop1,rhop2),0.5f);
        const float pi_visc=FMUL(FMUL(FMUL(FMUL(-visco,c.cs0),amubar),b2s_rcp(robar)),massp2);
        a.ax=FFMA(-pi_visc,frx,a.ax); a.ay=FFMA(-pi_visc,fry,a.ay); a.az=FFMA(-pi_visc,frz,a.az);
      }
    }
    else{
      const float robar2=FADD(rhop1,rhop2);
      const float temp=FMUL(FMUL(4.f,visco),b2s_rcp(FMUL(FADD(rr2,c.eta2),robar2)));
      const float vtemp=FMUL(FMUL(massp2,temp),dot3);
      a.ax=FFMA(vtemp,dvx,a.ax); a.ay=FFMA(vtemp,dvy,a.ay); a.az=FFMA(vtemp,dvz,a.az);
    }
    if(TVISCO==B2S_VISCO_LAMINARSPS){
      float sxx=taup1[0],sxy=taup1[1],sxz=taup1[2],syy=taup1[3],syz=taup1[4],szz=taup1[5];
      if(!BOUNDP2 && !ftp2){
        const float2* t2=(const float2*)(taug+size_t(p2)*6);
        const float2 t0=t2[0],t1=t2[1],t3=t2[2];
        sxx=FADD(sxx,t0.x); sxy=FADD(sxy,t0.y); sxz=FADD(sxz,t1.x); syy=FADD(syy,t1.y); syz=FADD(syz,t3.x); szz=FADD(szz,t3.y);
      }
      a.ax=FFMA(massp2,b2s_dot3(sxx,sxy,sxz,frx,fry,frz),a.ax);
      a.ay=FFMA(massp2,b2s_dot3(sxy,syy,syz,frx,fry,frz),a.ay);
      a.az=FFMA(massp2,b2s_dot3(sxz,syz,szz,frx,fry,frz),a.az);
      if(!ftp1){
        const float volp2=-FMUL(massp2,ovrhop2);
        float dv=FMUL(dvx,volp2); a.sxx=FFMA(dv,frx,a.sxx); a.sxy=FFMA(dv,fry,a.sxy); a.sxz=FFMA(dv,frz,a.sxz);
              dv=FMUL(dvy,volp2); a.sxy=FFMA(dv,frx,a.sxy); a.syy=FFMA(dv,fry,a.syy); a.syz=FFMA(dv,frz,a.syz);
              dv=FMUL(dvz,volp2); a.sxz=FFMA(dv,frx,a.sxz); a.syz=FFMA(dv,fry,a.syz); a.szz=FFMA(dv,frz,a.szz);
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

// Trims a row of neighbour cells (fixed y,z) to the cells [xa,xb) the kernel sphere of the target can reach.
// sy,sz = target y,z in the frame of the row's cells (pp1.y+offy): a neighbour there has y in [0,scell), so its distance
// in y is at least max(0,sy-scell,-sy), same in z; what is left of (2h)^2 bounds |dx|.  The 2e-4 margins keep the trim
// conservative: it only drops cells whose particles would all fail rr2<=4h^2, so results do not depend on it.
// xin = target x inside its own cell.  Returns false when the whole row is out of reach.
__device__ __forceinline__ bool row_trim(const DevCte& c,const CellRange& r,float xin,float sy,float sz,int& xa,int& xb){
  const float dy=fmaxf(0.f,fmaxf(sy-c.scell,-sy)), dz=fmaxf(0.f,fmaxf(sz-c.scell,-sz));
  const float xh2=c.kernelsize2*1.0002f-dy*dy-dz*dz;
  if(xh2<=0.f)return false;
  const float xh=xh2*b2s_rsqrt(xh2)+c.scell*1e-4f, ovs=1.0f/c.scell;   // approximate sqrt (2 ulp) is covered by the margins
  xa=max(r.xini,r.cx+int(floorf((xin-xh)*ovs)));
  xb=min(r.xfin,r.cx+int(floorf((xin+xh)*ovs))+1);
  return xa<xb;
}

// Store of one fluid/floating particle after both passes + SPS tau + the values for the fused maxima.
// (JSphCpu.cpp:1086-1102 fluid pass then bound pass; JSphGpu_ker.cu:939-950.)
// The density-diffusion sum is kept apart and added last, and a DBC boundary neighbour (FLT_MAX marker) discards it
// completely: this is what both reference builds do (GPU: one running deltap1, JSphGpu_ker.cu:939-946; CPU:
// DDTArray=true on CPU, JSph.cpp:927, Delta_c added after both passes, JSphCpuSingle.cpp:723-731).
// A pass merges its delta (and its FLT_MAX marker) only when it stores anything at all, i.e. under the reference's
// `if(arp1||acep1.x||acep1.y||acep1.z||visc)` (JSphCpu.cpp:1087): a free-surface particle at rest next to a DBC
// wall (P1=P2=0, dv=0) keeps its fluid-pass delta because the boundary pass stores nothing.
template<int TVISCO,int TDENS>
__device__ __forceinline__ void finalize_store(const ForcesArgs& g,const DevCte& c,uint32_t p1,const float4 velrhop1,
  const PairAcc& af,const PairAcc& ab,float& ace2,float& viscmax)
{
  float4 res;
  res.x=af.ax+ab.ax; res.y=af.ay+ab.ay; res.z=af.az+ab.az; res.w=af.ar+ab.ar;
  if(TDENS!=B2S_DDT_NONE){
    // with shifting the reference stores unconditionally: `if(shift||arp1||acep1.x||...)` (JSphCpu.cpp:1084)
    const bool sf=(c.shiftmode!=0 || af.ar!=0.f || af.ax!=0.f || af.ay!=0.f || af.az!=0.f || af.visc!=0.f);
    const bool sb=(c.shiftmode!=0 || ab.ar!=0.f || ab.ax!=0.f || ab.ay!=0.f || ab.az!=0.f || ab.visc!=0.f);
    const bool dmax=((sf && af.delta==FLT_MAX) || (sb && ab.delta==FLT_MAX));
    if(!dmax)res.w+=((sf? af.delta: 0.f)+(sb? ab.delta: 0.f));
  }
  if(c.simulate2d)res.y=0.f;
  g.acear[p1]=res;
  const bool owned=B2S_CODE_ISNORMAL(g.code[p1]);   // slab halo ghosts carry CODE_PERIODIC and do not enter the dt maxima
  viscmax=(owned? fmaxf(af.visc,ab.visc): 0.f);
  if(owned)ace2=res.x*res.x+res.y*res.y+res.z*res.z;
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

// One fluid/floating particle against all its neighbours, from global memory, in the reference's order: fluid neighbours, then
// boundary neighbours, rows of cells z-major (JSphCpu.cpp:736-1125; the walk of KerInteractionForcesFluid, JSphGpu_ker.cu:786-990).
// This is variant 0's thread body (forces.cu) and the fallback of the tile kernel for columns that do not fit its shared memory.
template<int TKER,int TVISCO,int TDENS,bool FLOATING>
__device__ __forceinline__ void gather_particle(const ForcesArgs& g,const DevCte& c,const DivState* ds,uint32_t p1,bool warpnearb,float& ace2,float& viscmax)
{
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
  const int ncx=int(ds->ncx), nsheet=int(ds->nsheet), cmx=int(ds->cellmin[0]);
  const float xin=pp1.x-c.scell*float((r.cx+cmx)&((1<<B2S_XBLK)-1));   // x inside the own cell (row_trim only)
  PairAcc af={0,0,0,0,0,0,0,0,0,0,0,0};
  PairAcc ab={0,0,0,0,0,0,0,0,0,0,0,0};
  if(FLOATING && ftp1 && TDENS!=B2S_DDT_NONE){ af.delta=FLT_MAX; ab.delta=FLT_MAX; }
  //-Fluid/floating neighbours.
  {
    const uint32_t* beg=g.beg+ds->box_fluid;
    for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
      const int row=y*ncx+z*nsheet;
      const float offy=c.scell*float(r.cy-y), offz=c.scell*float(r.cz-z);
      int xa,xb;
      if(!row_trim(c,r,xin,FADD(pp1.y,offy),FADD(pp1.z,offz),xa,xb))continue;
      for(int x=xa;x<xb;x++){
        const uint32_t pini=beg[row+x], pfin=beg[row+x+1];
        const float offx=b2s_offx(c,r.cx+cmx,x+cmx);
        for(uint32_t p2=pini;p2<pfin;p2++){
          const float4 pp2=g.posp[p2];
          const float drx=b2s_dr(pp1.x,pp2.x,offx), dry=b2s_dr(pp1.y,pp2.y,offy), drz=b2s_dr(pp1.z,pp2.z,offz);
          const float rr2=b2s_rr2(drx,dry,drz);
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
  if(warpnearb){
    const uint32_t* beg=g.beg;
    for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
      const int row=y*ncx+z*nsheet;
      const float offy=c.scell*float(r.cy-y), offz=c.scell*float(r.cz-z);
      int xa,xb;
      if(!row_trim(c,r,xin,FADD(pp1.y,offy),FADD(pp1.z,offz),xa,xb))continue;
      for(int x=xa;x<xb;x++){
        const uint32_t pini=beg[row+x], pfin=beg[row+x+1];
        const float offx=b2s_offx(c,r.cx+cmx,x+cmx);
        for(uint32_t p2=pini;p2<pfin;p2++){
          const float4 pp2=g.posp[p2];
          const float drx=b2s_dr(pp1.x,pp2.x,offx), dry=b2s_dr(pp1.y,pp2.y,offy), drz=b2s_dr(pp1.z,pp2.z,offz);
          const float rr2=b2s_rr2(drx,dry,drz);
          if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO){
            pair_fluid<TKER,TVISCO,TDENS,FLOATING,true>(c,drx,dry,drz,rr2,velrhop1,pressp1,ftp1,taup1,
              g.velrho[p2],pp2.w,typecode(0),g.tau,p2,g.ftomassp,c.visco_b,ab,g.boundmode,g.tangenvel);
          }
        }
      }
    }
  }
  finalize_store<TVISCO,TDENS>(g,c,p1,velrhop1,af,ab,ace2,viscmax);
}

// Warp max of the two dt quantities + one atomic per warp (replaces ReduMaxFloat x2 + cudaMemcpy, JSphGpu_ker.cu:75-125).
__device__ __forceinline__ void reduce_maxima(StepState* st,float viscmax,float ace2){
  #pragma unroll
  for(int o=16;o;o>>=1){ viscmax=fmaxf(viscmax,__shfl_xor_sync(0xffffffffu,viscmax,o)); ace2=fmaxf(ace2,__shfl_xor_sync(0xffffffffu,ace2,o)); }
  if((threadIdx.x&31)==0){
    if(viscmax>0.f)atomicMax(&st->viscdtmax_bits,__float_as_uint(viscmax));
    if(ace2>0.f)atomicMax(&st->acemax2_bits,__float_as_uint(ace2));
  }
}
