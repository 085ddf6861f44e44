// forces_pair2.cuh -- packed-FP32 (f32x2) helpers and the two-neighbours-per-lane pair body shared by the packed interaction
// kernels (forces_pair2.cu: variants 4/5, forces_tile.cu: variant 6).  Physics and operation order: pair_fluid (forces_common.cuh),
// i.e. JSphCpu.cpp:822-1041 / JSphGpu_ker.cu:590-774.
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

// Pair-loop constants derived from DevCte once per launch (host side).
struct PackCte { float mhalf_ovh, cvisc_f, cvisc_b, ddtk; };
static inline PackCte pack_cte(const DevCte& c){
  PackCte k; k.mhalf_ovh=-0.5f*c.ovkernelh; k.cvisc_f=-2.f*c.visco_f*c.cs0*c.kernelh; k.cvisc_b=-2.f*c.visco_b*c.cs0*c.kernelh; k.ddtk=c.ddtkh*c.cs0;
  return k;
}

// Unscaled sums of one pass (fluid or boundary neighbours) of one target: the factors common to every pair of the pass are applied
// once at the end (PackAcc -> PairAcc in pack_finish).
struct PackAcc { f2 ax,ay,az,ar,delta; float visc; bool mark; };

// Two hits (a,b) of target p1.  Same terms as pair_fluid (forces_common.cuh; JSphCpu.cpp:822-1041, JSphGpu_ker.cu:590-774) for
// Wendland + artificial viscosity, re-associated so that the kernel-gradient VECTOR fac*dr is never formed:
//   fac = bwenh*w^3, w = 1-q/2                                       (FunSphKernel.h:218-223)
//   ace_i += -m*fac*(A+PI)*dr,  A = (Pi+Pj)/(rhoi*rhoj),  PI = -visco*cs0*h*min(dot,0)/((rr2+eta2)*(rhoi+rhoj)/2),  dot = dr.dv
//   ar_i  += m*fac*dot*rhoi/rhoj                                      (dv.(fac*dr) = fac*dot)
//   delta -= ddtkh*cs0*((rhoj-rhoi)-drho)/(rr2+eta2) * fac*rr2 * m/rhoj     (dr.(fac*dr) = fac*rr2)
// and the pair-independent factors (-m*bwenh, m*bwenh*rhoi, ddtkh*cs0*m*bwenh) wait for pack_finish.  Against the reference's
// order of operations this changes the rounding of single terms by ~1e-7 relative (tests/test_gpu_parity.py states the bound).
// A pair at distance 0 -- the target itself, which is not excluded from its list and also pads an odd hit count, or a coincident
// particle, which the reference skips below rr2=1e-18 (JSphGpu_ker.cu:563) -- contributes exactly 0: dr=0, and the 1e-30 under
// the rsqrt keeps w finite.
template<int TDENS,bool BOUNDP2>
__device__ __forceinline__ void pack_body(const DevCte& c,const PackCte& k,f2 drx,f2 dry,f2 drz,f2 rr2,
  const float4 velrhop1,float pressp1,float ovrhop1,const float4 va,const float4 vb,float pressa,float pressb,PackAcc& a)
{
  const float rhop1=velrhop1.w;
  const f2 dvx=pk(FSUB(velrhop1.x,va.x),FSUB(velrhop1.x,vb.x)), dvy=pk(FSUB(velrhop1.y,va.y),FSUB(velrhop1.y,vb.y)), dvz=pk(FSUB(velrhop1.z,va.z),FSUB(velrhop1.z,vb.z));
  const f2 dot=dot3_2(drx,dry,drz,dvx,dvy,dvz);
  const f2 rg=add2(rr2,pk1(1e-30f));
  const f2 rad=mul2(rr2,pk(b2s_rsqrt(lo(rg)),b2s_rsqrt(hi(rg))));
  f2 w1=fma2(rad,pk1(k.mhalf_ovh),pk1(1.f));
  w1=pk(fmaxf(lo(w1),0.f),fmaxf(hi(w1),0.f));                        // beyond 2h (false positives of the FP16 pre-filter): no contribution
  const f2 w3=mul2(mul2(w1,w1),w1);                                  // fac/bwenh
  const f2 ovrhop2=pk(b2s_rcp(va.w),b2s_rcp(vb.w));
  const f2 rr2eta=add2(rr2,pk1(c.eta2));
  const f2 ovr=pk(b2s_rcp(lo(rr2eta)),b2s_rcp(hi(rr2eta)));
  const f2 dotr=mul2(dot,ovr);
  a.visc=fmaxf(a.visc,fmaxf(lo(dotr),hi(dotr)));
  //-Momentum: pressure + artificial viscosity (only approaching pairs: min(dot,0)).
  const f2 A=mul2(pk(FADD(pressp1,pressa),FADD(pressp1,pressb)),mul2(pk1(ovrhop1),ovrhop2));
  const f2 rsum=pk(FADD(rhop1,va.w),FADD(rhop1,vb.w));
  const f2 pin=mul2(pk(fminf(lo(dotr),0.f),fminf(hi(dotr),0.f)),pk(b2s_rcp(lo(rsum)),b2s_rcp(hi(rsum))));
  const f2 S=fma2(pk1(BOUNDP2? k.cvisc_b: k.cvisc_f),pin,A);
  const f2 coef=mul2(S,w3);
  a.ax=fma2(coef,drx,a.ax); a.ay=fma2(coef,dry,a.ay); a.az=fma2(coef,drz,a.az);
  //-Continuity.
  a.ar=fma2(mul2(w3,dot),ovrhop2,a.ar);
  //-Density diffusion.
  if(TDENS==B2S_DDT_MOLTENI){
    if(BOUNDP2 && c.tboundary==B2S_BC_DBC)a.mark=true;
    else a.delta=fma2(mul2(fma2(pk1(rhop1),ovrhop2,pk1(-1.f)),ovr),mul2(w3,rr2),a.delta);
  }
  if(TDENS==B2S_DDT_FOURTAKAS || (TDENS==B2S_DDT_FOURTAKASFULL && !BOUNDP2)){
    if(BOUNDP2)a.mark=true;
    else{
      const f2 one=pk1(1.f);
      const f2 x=sub2(fma2(pk1(c.ddtgz),drz,one),one);             // fl32(1+ddtgz*drz)-1: the rounding of the reference's pow() argument
      const f2 pw=fma2(x,fma2(x,fma2(x,pk1(c.ddt_c3),pk1(c.ddt_c2)),pk1(c.ddt_c1)),one);
      const f2 rho0=pk1(c.rhopzero);
      const f2 drho=sub2(mul2(rho0,pw),rho0);
      const f2 dd=sub2(pk(FSUB(va.w,rhop1),FSUB(vb.w,rhop1)),drho);
      a.delta=fma2(mul2(mul2(dd,ovr),mul2(w3,rr2)),ovrhop2,a.delta);
    }
  }
}

// Scales the sums of one pass and hands them over in the form finalize_store expects.
template<int TDENS,bool BOUNDP2>
__device__ __forceinline__ void pack_finish(const DevCte& c,const PackCte& k,const PackAcc& a,float rhop1,PairAcc& o){
  const float m=(BOUNDP2? c.massb: c.massf);
  const float fa=-m*c.bwenh, fr=m*c.bwenh*rhop1, fd=k.ddtk*m*c.bwenh;
  o.ax=fa*(lo(a.ax)+hi(a.ax)); o.ay=fa*(lo(a.ay)+hi(a.ay)); o.az=fa*(lo(a.az)+hi(a.az));
  o.ar=fr*(lo(a.ar)+hi(a.ar));
  o.visc=a.visc;
  // Molteni adds, Fourtakas subtracts (JSphCpu.cpp:934-950)
  o.delta=(a.mark? FLT_MAX: (TDENS==B2S_DDT_MOLTENI? fd: -fd)*(lo(a.delta)+hi(a.delta)));
  if(TDENS==B2S_DDT_NONE)o.delta=0.f;
}

