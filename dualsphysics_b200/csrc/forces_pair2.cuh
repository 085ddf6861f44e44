// forces_pair2.cuh -- packed-FP32 (f32x2) helpers and the two-neighbours-per-lane pair body shared by the packed interaction
// kernels (forces_pair2.cu: variant 5, forces_tile.cu: variant 6).  Physics and operation order: pair_fluid (forces_common.cuh),
// i.e. JSphCpu.cpp:822-1041 / JSphGpu_ker.cu:590-774, for every template axis of the reference's dispatch
// (JSphGpu_ker.cu:1080-1179): Wendland / cubic spline, artificial / laminar / laminar+SPS viscosity, every DDT mode, floating
// bodies (ftomassp) and the mDBC2 boundary modes.
#pragma once
#include "forces_common.cuh"

// ---- f32x2 helpers (PTX ISA 8.6+, sm_100): a 64-bit register holds (lo,hi) ----------------------------------------------
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo,float hi){ f2 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(lo),"f"(hi)); return r; }
__device__ __forceinline__ f2 pk1(float v){ return pk(v,v); }
__device__ __forceinline__ float lo(f2 a){ float x,y; asm("mov.b64 {%0,%1}, %2;":"=f"(x),"=f"(y):"l"(a)); return x; }
__device__ __forceinline__ float hi(f2 a){ float x,y; asm("mov.b64 {%0,%1}, %2;":"=f"(x),"=f"(y):"l"(a)); return y; }
__device__ __forceinline__ f2 mul2(f2 a,f2 b){ f2 d; asm("mul.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d; }
__device__ __forceinline__ f2 add2(f2 a,f2 b){ f2 d; asm("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a,f2 b){ f2 d; asm("sub.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d; }
__device__ __forceinline__ f2 fma2(f2 a,f2 b,f2 c){ f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d; }
__device__ __forceinline__ f2 dot3_2(f2 ax,f2 ay,f2 az,f2 bx,f2 by,f2 bz){ return fma2(az,bz,fma2(ay,by,mul2(ax,bx))); }   // b2s_dot3 order
__device__ __forceinline__ f2 rcp2(f2 a){ return pk(b2s_rcp(lo(a)),b2s_rcp(hi(a))); }
__device__ __forceinline__ f2 sel2(bool plo,bool phi,f2 a,f2 b){ return pk(plo? lo(a): lo(b),phi? hi(a): hi(b)); }

// Pair-loop constants derived from DevCte once per launch (host side).
struct PackCte { float mhalf_ovh, cvisc_f, cvisc_b, ddtk, lam_f, lam_b, ovmassf; };
static inline PackCte pack_cte(const DevCte& c){
  PackCte k; k.mhalf_ovh=-0.5f*c.ovkernelh; k.cvisc_f=-2.f*c.visco_f*c.cs0*c.kernelh; k.cvisc_b=-2.f*c.visco_b*c.cs0*c.kernelh; k.ddtk=c.ddtkh*c.cs0;
  k.lam_f=-4.f*c.visco_f; k.lam_b=-4.f*c.visco_b; k.ovmassf=1.0f/c.massf;
  return k;
}

// Unscaled sums of one pass (fluid or boundary neighbours) of one target: the factors common to every pair of the pass are applied
// once at the end (PackAcc -> PairAcc in pack_finish).  s?? = the 2*strain sums of laminar+SPS (dead code otherwise).
struct PackAcc { f2 ax,ay,az,ar,delta; f2 sxx,sxy,sxz,syy,syz,szz; float visc; bool mark; };
__device__ __forceinline__ void pack_zero(PackAcc& a){
  a.ax=a.ay=a.az=a.ar=a.delta=0ull; a.sxx=a.sxy=a.sxz=a.syy=a.syz=a.szz=0ull; a.visc=0.f; a.mark=false;
}

// What a round knows about its target beyond position and velocity (all dead code for Wendland + artificial without floating).
struct PackTarget {
  float tensil;      // cubic: P/rho^2 * (P>0? 0.01: -0.2)   (FunSphKernel.h:88-97)
  float tau[6];      // laminar+SPS: own stress tensor
  bool ft;           // floating particle
};
// ... and about its two neighbours beyond posp / velrho.
struct PackNeigh {
  f2 mu;             // mass of the neighbour relative to the mass of the pass (1 unless floating or an mDBC2 particle without mass)
  f2 plain;          // 1 = plain fluid neighbour (its tau counts, the Fourtakas term counts), 0 = floating
  bool heavy_a,heavy_b;   // floating neighbour lighter than 1.2 fluid masses: DELTA_HEAVYFLOATING marker (DualSphDef.h:104)
  const float2 *taua,*taub;   // laminar+SPS, fluid pass
  float4 tva,tvb;             // mDBC2: the viscous terms see the tangential velocity of the boundary particle (JSphCpu.cpp:990-995)
};

// Two hits (a,b) of target p1.  Same terms as pair_fluid (forces_common.cuh; JSphCpu.cpp:822-1041, JSphGpu_ker.cu:590-774),
// re-associated so that the kernel-gradient VECTOR fac*dr is never formed.  With facn = fac/K (K = bwenh for Wendland, where
// facn = w^3 with w = 1-q/2, FunSphKernel.h:218-223; K = 1 for the cubic spline) and mu = m_j/m:
//   ace_i += -m*K * facn*mu*(A+PI)*dr,  A = (Pi+Pj)/(rhoi*rhoj) [+ tensile term],  PI = -visco*cs0*h*min(dot,0)/((rr2+eta2)*(rhoi+rhoj)/2)
//   laminar instead of PI:  ace_i += m*K * 4*visco*facn*mu*rr2/((rr2+eta2)*(rhoi+rhoj)) * dv      (dr.(fac*dr) = fac*rr2)
//   SPS:    ace_i += m*K * facn*mu * (tau_i+tau_j).dr ;   2*strain_i += -m*K * facn*mu/rhoj * (dv (x) dr)
//   ar_i  += m*K*rhoi * facn*mu*dot/rhoj,  dot = dr.dv                                             (dv.(fac*dr) = fac*dot)
//   delta -= ddtkh*cs0*m*K * ((rhoj-rhoi)-drho)/(rr2+eta2) * facn*mu*rr2/rhoj
// and the pair-independent factors wait for pack_finish.  Against the reference's order of operations this changes the rounding
// of single terms by ~1e-7 relative (tests/test_gpu_parity.py states the bound).
// A pair at distance 0 -- the target itself, which is not excluded from its list and also pads an odd hit count, or a coincident
// particle, which the reference skips below rr2=1e-18 (JSphGpu_ker.cu:563) -- contributes exactly 0: dr=0, and the 1e-30 under
// the rsqrt keeps the kernel factor finite.  A false positive of the FP16 pre-filter (rr2 > (2h)^2) contributes 0 as well: the
// kernel factor is clamped (Wendland) or selected (cubic) to zero, and the FLT_MAX markers test rr2 exactly.
template<int TKER,int TVISCO,int TDENS,bool FLOATING,bool MDBC2,bool BOUNDP2>
__device__ __forceinline__ void pack_body(const DevCte& c,const PackCte& k,f2 drx,f2 dry,f2 drz,f2 rr2,
  const float4 velrhop1,float pressp1,float ovrhop1,const PackTarget& t,
  const float4 va,const float4 vb,float pressa,float pressb,const PackNeigh& n,PackAcc& a)
{
  const float rhop1=velrhop1.w;
  const bool vmass=((FLOATING && !BOUNDP2) || (BOUNDP2 && MDBC2));      // neighbour masses differ inside the pass
  const f2 dvx=pk(FSUB(velrhop1.x,va.x),FSUB(velrhop1.x,vb.x)), dvy=pk(FSUB(velrhop1.y,va.y),FSUB(velrhop1.y,vb.y)), dvz=pk(FSUB(velrhop1.z,va.z),FSUB(velrhop1.z,vb.z));
  const f2 dot=dot3_2(drx,dry,drz,dvx,dvy,dvz);
  //-Kernel factor facn.
  f2 facn, tens=0ull;
  if(TKER==B2S_KERNEL_WENDLAND){
    const f2 rg=add2(rr2,pk1(1e-30f));
    const f2 rad=mul2(rr2,pk(b2s_rsqrt(lo(rg)),b2s_rsqrt(hi(rg))));
    f2 w1=fma2(rad,pk1(k.mhalf_ovh),pk1(1.f));
    w1=pk(fmaxf(lo(w1),0.f),fmaxf(hi(w1),0.f));                      // beyond 2h (false positives of the FP16 pre-filter): no contribution
    facn=mul2(mul2(w1,w1),w1);                                       // fac/bwenh
  }
  else{
    // cubic spline (FunSphKernel.h:53-65) and its tensile correction (:88-97), both branches evaluated and selected per half
    const bool ina=(lo(rr2)<=c.kernelsize2), inb=(hi(rr2)<=c.kernelsize2);
    const f2 rc=pk(fminf(lo(rr2),4.f*c.kernelsize2),fminf(hi(rr2),4.f*c.kernelsize2));   // keeps everything finite for the padded hit
    const f2 rg=add2(rc,pk1(1e-30f));
    const f2 ovrad=pk(b2s_rsqrt(lo(rg)),b2s_rsqrt(hi(rg)));
    const f2 rad=mul2(rc,ovrad);
    const f2 qq=mul2(rad,pk1(c.ovkernelh));
    const bool fara=(lo(rad)>c.kernelh), farb=(hi(rad)>c.kernelh);
    const f2 wq1=sub2(pk1(2.f),qq), wq2f=mul2(wq1,wq1), wq2n=mul2(qq,qq);
    const f2 facf=mul2(mul2(pk1(c.cub_c2),wq2f),ovrad);
    const f2 facc=mul2(fma2(pk1(c.cub_d1),wq2n,mul2(pk1(c.cub_c1),qq)),ovrad);
    facn=sel2(ina,inb,sel2(fara,farb,facf,facc),0ull);
    const f2 wabf=mul2(pk1(c.cub_a24),mul2(wq2f,wq1));
    const f2 wabn=mul2(pk1(c.cub_a2),fma2(fma2(pk1(0.75f),qq,pk1(-1.5f)),wq2n,pk1(1.f)));
    f2 fab=mul2(sel2(fara,farb,wabf,wabn),pk1(c.cub_od_wdeltap));
    fab=mul2(fab,fab); fab=mul2(fab,fab);
    const f2 ov2=pk(b2s_rcp(FMUL(va.w,va.w)),b2s_rcp(FMUL(vb.w,vb.w)));
    const f2 t2=mul2(mul2(pk(pressa,pressb),ov2),pk(pressa>0.f? 0.01f: -0.2f,pressb>0.f? 0.01f: -0.2f));
    tens=mul2(fab,add2(pk1(t.tensil),t2));
  }
  const f2 ovrhop2=pk(b2s_rcp(va.w),b2s_rcp(vb.w));
  const f2 rr2eta=add2(rr2,pk1(c.eta2));
  const f2 ovr=rcp2(rr2eta);
  //-Velocity difference of the viscous terms.
  f2 wvx=dvx, wvy=dvy, wvz=dvz, dotv=dot;
  if(BOUNDP2 && MDBC2){
    wvx=pk(FSUB(velrhop1.x,n.tva.x),FSUB(velrhop1.x,n.tvb.x)); wvy=pk(FSUB(velrhop1.y,n.tva.y),FSUB(velrhop1.y,n.tvb.y)); wvz=pk(FSUB(velrhop1.z,n.tva.z),FSUB(velrhop1.z,n.tvb.z));
    dotv=dot3_2(drx,dry,drz,wvx,wvy,wvz);
  }
  const f2 dotr=mul2(dotv,ovr);
  a.visc=fmaxf(a.visc,fmaxf(lo(dotr),hi(dotr)));
  const f2 facm=(vmass? mul2(facn,n.mu): facn);
  //-Momentum: pressure (+ tensile correction) and viscosity.
  f2 A=mul2(pk(FADD(pressp1,pressa),FADD(pressp1,pressb)),mul2(pk1(ovrhop1),ovrhop2));
  if(TKER==B2S_KERNEL_CUBIC)A=add2(A,tens);
  const f2 rsum=pk(FADD(rhop1,va.w),FADD(rhop1,vb.w));
  if(TVISCO==B2S_VISCO_ARTIFICIAL){
    //-only approaching pairs: min(dot,0)
    const f2 pin=mul2(pk(fminf(lo(dotr),0.f),fminf(hi(dotr),0.f)),rcp2(rsum));
    const f2 coef=mul2(fma2(pk1(BOUNDP2? k.cvisc_b: k.cvisc_f),pin,A),facm);
    a.ax=fma2(coef,drx,a.ax); a.ay=fma2(coef,dry,a.ay); a.az=fma2(coef,drz,a.az);
  }
  else{
    const f2 coef=mul2(A,facm);
    const f2 lam=mul2(mul2(mul2(facm,rr2),pk1(BOUNDP2? k.lam_b: k.lam_f)),rcp2(mul2(rr2eta,rsum)));
    a.ax=fma2(coef,drx,fma2(lam,wvx,a.ax)); a.ay=fma2(coef,dry,fma2(lam,wvy,a.ay)); a.az=fma2(coef,drz,fma2(lam,wvz,a.az));
  }
  if(TVISCO==B2S_VISCO_LAMINARSPS){
    f2 sxx=pk1(t.tau[0]),sxy=pk1(t.tau[1]),sxz=pk1(t.tau[2]),syy=pk1(t.tau[3]),syz=pk1(t.tau[4]),szz=pk1(t.tau[5]);
    if(!BOUNDP2){
      const float2 a0=__ldg(n.taua),a1=__ldg(n.taua+1),a2=__ldg(n.taua+2), b0=__ldg(n.taub),b1=__ldg(n.taub+1),b2=__ldg(n.taub+2);
      if(FLOATING){   // the stress of a floating neighbour does not count
        sxx=fma2(n.plain,pk(a0.x,b0.x),sxx); sxy=fma2(n.plain,pk(a0.y,b0.y),sxy); sxz=fma2(n.plain,pk(a1.x,b1.x),sxz);
        syy=fma2(n.plain,pk(a1.y,b1.y),syy); syz=fma2(n.plain,pk(a2.x,b2.x),syz); szz=fma2(n.plain,pk(a2.y,b2.y),szz);
      }
      else{
        sxx=add2(sxx,pk(a0.x,b0.x)); sxy=add2(sxy,pk(a0.y,b0.y)); sxz=add2(sxz,pk(a1.x,b1.x));
        syy=add2(syy,pk(a1.y,b1.y)); syz=add2(syz,pk(a2.x,b2.x)); szz=add2(szz,pk(a2.y,b2.y));
      }
    }
    const f2 nf=sub2(0ull,facm);
    a.ax=fma2(nf,dot3_2(sxx,sxy,sxz,drx,dry,drz),a.ax);
    a.ay=fma2(nf,dot3_2(sxy,syy,syz,drx,dry,drz),a.ay);
    a.az=fma2(nf,dot3_2(sxz,syz,szz,drx,dry,drz),a.az);
    if(!(FLOATING && t.ft)){
      const f2 g=mul2(facm,ovrhop2);
      f2 dv=mul2(wvx,g); a.sxx=fma2(dv,drx,a.sxx); a.sxy=fma2(dv,dry,a.sxy); a.sxz=fma2(dv,drz,a.sxz);
         dv=mul2(wvy,g); a.sxy=fma2(dv,drx,a.sxy); a.syy=fma2(dv,dry,a.syy); a.syz=fma2(dv,drz,a.syz);
         dv=mul2(wvz,g); a.sxz=fma2(dv,drx,a.sxz); a.syz=fma2(dv,dry,a.syz); a.szz=fma2(dv,drz,a.szz);
    }
  }
  //-Continuity.
  a.ar=fma2(mul2(facm,dot),ovrhop2,a.ar);
  //-Density diffusion.  The FLT_MAX markers of the reference sit inside its rr2<=(2h)^2 test: exact test here.
  const bool hit=(lo(rr2)<=c.kernelsize2 || hi(rr2)<=c.kernelsize2);
  if(TDENS==B2S_DDT_MOLTENI){
    if(FLOATING && !BOUNDP2 && ((n.heavy_a && lo(rr2)<=c.kernelsize2 && lo(rr2)>=B2S_ALMOSTZERO) || (n.heavy_b && hi(rr2)<=c.kernelsize2 && hi(rr2)>=B2S_ALMOSTZERO)))a.mark=true;
    if(BOUNDP2 && c.tboundary==B2S_BC_DBC){ if(hit)a.mark=true; }
    else a.delta=fma2(mul2(fma2(pk1(rhop1),ovrhop2,pk1(-1.f)),ovr),mul2(facm,rr2),a.delta);
  }
  if(TDENS==B2S_DDT_FOURTAKAS || (TDENS==B2S_DDT_FOURTAKASFULL && !BOUNDP2)){
    if(BOUNDP2){ if(hit)a.mark=true; }
    else{
      const f2 one=pk1(1.f);
      const f2 x=sub2(fma2(pk1(c.ddtgz),drz,one),one);             // fl32(1+ddtgz*drz)-1: the rounding of the reference's pow() argument
      const f2 pw=fma2(x,fma2(x,fma2(x,pk1(c.ddt_c3),pk1(c.ddt_c2)),pk1(c.ddt_c1)),one);
      const f2 rho0=pk1(c.rhopzero);
      const f2 drho=sub2(mul2(rho0,pw),rho0);
      const f2 dd=sub2(pk(FSUB(va.w,rhop1),FSUB(vb.w,rhop1)),drho);
      f2 term=mul2(mul2(dd,ovr),mul2(facm,rr2));
      if(FLOATING)term=mul2(term,n.plain);                         // `&& !ftp2` (JSphCpu.cpp:943)
      a.delta=fma2(term,ovrhop2,a.delta);
    }
  }
}

// Scales the sums of one pass and hands them over in the form finalize_store expects.
template<int TKER,int TVISCO,int TDENS,bool BOUNDP2>
__device__ __forceinline__ void pack_finish(const DevCte& c,const PackCte& k,const PackAcc& a,float rhop1,bool deltaoff,PairAcc& o){
  const float m=(BOUNDP2? c.massb: c.massf);
  const float mk=(TKER==B2S_KERNEL_WENDLAND? m*c.bwenh: m);
  const float fa=-mk, fr=mk*rhop1, fd=k.ddtk*mk;
  o.ax=fa*(lo(a.ax)+hi(a.ax)); o.ay=fa*(lo(a.ay)+hi(a.ay)); o.az=fa*(lo(a.az)+hi(a.az));
  o.ar=fr*(lo(a.ar)+hi(a.ar));
  o.visc=a.visc;
  // Molteni adds, Fourtakas subtracts (JSphCpu.cpp:934-950); deltaoff: a floating target never takes part (JSphCpu.cpp:811)
  o.delta=((a.mark || deltaoff)? FLT_MAX: (TDENS==B2S_DDT_MOLTENI? fd: -fd)*(lo(a.delta)+hi(a.delta)));
  if(TDENS==B2S_DDT_NONE)o.delta=0.f;
  if(TVISCO==B2S_VISCO_LAMINARSPS){
    o.sxx=fa*(lo(a.sxx)+hi(a.sxx)); o.sxy=fa*(lo(a.sxy)+hi(a.sxy)); o.sxz=fa*(lo(a.sxz)+hi(a.sxz));
    o.syy=fa*(lo(a.syy)+hi(a.syy)); o.syz=fa*(lo(a.syz)+hi(a.syz)); o.szz=fa*(lo(a.szz)+hi(a.szz));
  }
}
