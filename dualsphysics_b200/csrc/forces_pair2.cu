// forces_pair2.cu -- interaction variant 4: two neighbours per lane and iteration, pair arithmetic in packed FP32 (sm_100a).
//
// Same physics as variant 0 (forces.cu / forces_common.cuh pair_fluid), which restates cusph::Interaction_Forces /
// KerInteractionForcesFluid (src/source/JSphGpu_ker.cu:673-964) and JSphCpu::InteractionForcesFluid (src/source/JSphCpu.cpp:736-1125).
//
// Why: variant 0 is bound by FP32 *issue slots* (profiles/README.md: 80% of the slots busy, FMA pipe 45%, 19 of 32 lanes active
// because the ~70-instruction pair body sits under the rr2<(2h)^2 branch).  Blackwell's fma/mul/add.rn.f32x2 do two FP32
// operations per issue slot (tools/microbench.cu: same FMA throughput as FFMA, half the slots).  Here every lane takes TWO
// consecutive candidates of its row per iteration; the body runs when either is a hit and evaluates both in f32x2 registers,
// the miss (if any) being neutralised by a 0/1 mask on the kernel gradient instead of a branch.  Per element the operations
// and their rounding are exactly those of pair_fluid; only the grouping of the sums differs (even/odd candidates accumulate
// in the two halves and are added at the end), so results agree with variant 0 to rounding, not bit for bit.
//   * rows of neighbour cells are walked as ONE index range (block-relative x, b2s_internal.h B2S_XBLK) and trimmed to the
//     cells the kernel sphere reaches (row_trim);
//   * built for Wendland + artificial viscosity without floating bodies (every DDT mode, DBC and mDBC); other
//     configurations run variant 0.
#include "forces_common.cuh"
#include <cuda_fp16.h>

#ifndef P2_BS
#define P2_BS 128
#endif
#ifndef P2_MINB
#define P2_MINB 7            // resident blocks per SM the register allocation aims at
#endif
#ifndef P2_MASK
#define P2_MASK 2            // 1: hits of a row chunk are collected in a bit mask (FP32 test), then evaluated two at a time;
                             // 2: same with the FP16 pre-filter on posh; 0: two consecutive candidates per iteration
#endif
#ifndef P2_PIPE
#define P2_PIPE 0            // 1: body rounds software-pipelined two deep (measured slower: 126 registers or heavy spills, profiles/r01_pair2_tuning.txt)
#endif
#ifndef P2_EARLYV
#define P2_EARLYV 0          // 1: load velrho of both candidates together with posp (no dependent load inside the body)
#endif

#ifndef P2_VNOW
#define P2_VNOW 1            // 1 (default since r02: 11.41 -> 10.91 ms at 10M, profiles/r02_pair2_ab.txt): the exact-test branch of a round also depends on the velrho gathers, so
                             // that ptxas cannot sink them below it and a round waits once for four gathers instead of twice for two
                             // (profiles/README.md: the wait at their first use holds 8.8% of the stall samples; a volatile load does
                             // not help, the sinking happens in ptxas).  Mirror-row loop only.
#endif

#ifndef P2_CHK2
#define P2_CHK2 1            // 1 (default since r02: 11.41 -> 10.55 ms, both together 10.16 ms): leaner FP16 pre-filter of the mirror-row loop -- no per-pair index clamp
                             // (posh is padded), mask bits as immediates, one x-block select per pair instead of two.
#endif

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

// FP16 pre-filter: rr2 of two candidates (half2) against the threshold; sets bit0/bit1 in mask for lo/hi
__device__ __forceinline__ void hit2_or(uint32_t& mask,uint32_t rr2,uint32_t thr2,uint32_t bit0,uint32_t bit1){
  asm("{\n .reg .pred p,q;\n setp.le.f16x2 p|q, %1, %2;\n @p or.b32 %0, %0, %3;\n @q or.b32 %0, %0, %4;\n}"
      :"+r"(mask):"r"(rr2),"r"(thr2),"r"(bit0),"r"(bit1));
}
__device__ __forceinline__ uint32_t h2u(__half2 h){ return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u){ return *reinterpret_cast<__half2*>(&u); }

struct Acc2 { f2 ax,ay,az,ar,delta; float visc; bool mark; };

// Two neighbours (a: p2a, b: p2b) of target p1.  ha/hb: which of them are real pairs.  Follows pair_fluid operation by operation.
template<int TDENS,bool BOUNDP2>
__device__ __forceinline__ void pair2_body(const DevCte& c,f2 drx,f2 dry,f2 drz,f2 rr2,bool ha,bool hb,
  const float4 velrhop1,float pressp1,const float4 va,const float4 vb,float pressa,float pressb,float visco,Acc2& a)
{
  const float massp2=(BOUNDP2? c.massb: c.massf);
  const f2 mass2=pk1(massp2);
  //-Misses are evaluated at rr2=(2h)^2 (finite everywhere) and removed by the 0/1 mask on the gradient.
  const f2 rr2s=pk(ha? lo(rr2): c.kernelsize2, hb? hi(rr2): c.kernelsize2);
  const f2 m2=pk(ha? 1.f: 0.f, hb? 1.f: 0.f);
  //-kernel_fac<WENDLAND>: rad=rr2*rsqrt(rr2); qq=rad/h; wqq1=1-0.5qq; fac=((bwenh*wqq1)*wqq1)*wqq1
  const f2 rad=mul2(rr2s,pk(b2s_rsqrt(lo(rr2s)),b2s_rsqrt(hi(rr2s))));
  const f2 qq=mul2(rad,pk1(c.ovkernelh));
  const f2 wqq1=fma2(pk1(-0.5f),qq,pk1(1.f));
  f2 fac=mul2(mul2(mul2(pk1(c.bwenh),wqq1),wqq1),wqq1);
  fac=mul2(fac,m2);
  const f2 frx=mul2(fac,drx), fry=mul2(fac,dry), frz=mul2(fac,drz);
  const float rhop1=velrhop1.w;
  const f2 ovrhop2=pk(b2s_rcp(va.w),b2s_rcp(vb.w));
  //-Momentum equation.
  {
    const f2 prs=mul2(pk(FADD(pressp1,pressa),FADD(pressp1,pressb)),mul2(pk1(b2s_rcp(rhop1)),ovrhop2));
    const f2 p_vpm=mul2(prs,pk1(-massp2));
    a.ax=fma2(p_vpm,frx,a.ax); a.ay=fma2(p_vpm,fry,a.ay); a.az=fma2(p_vpm,frz,a.az);
  }
  //-Continuity equation.
  const f2 dvx=pk(FSUB(velrhop1.x,va.x),FSUB(velrhop1.x,vb.x)), dvy=pk(FSUB(velrhop1.y,va.y),FSUB(velrhop1.y,vb.y)), dvz=pk(FSUB(velrhop1.z,va.z),FSUB(velrhop1.z,vb.z));
  a.ar=fma2(mul2(mass2,dot3_2(dvx,dvy,dvz,frx,fry,frz)),mul2(pk1(rhop1),ovrhop2),a.ar);
  const f2 rr2eta=add2(rr2s,pk1(c.eta2));
  const f2 ovrr2eta=pk(b2s_rcp(lo(rr2eta)),b2s_rcp(hi(rr2eta)));
  //-Density diffusion.
  if(TDENS==B2S_DDT_MOLTENI){
    if(BOUNDP2 && c.tboundary==B2S_BC_DBC)a.mark=true;
    else{
      const f2 dot3=dot3_2(drx,dry,drz,frx,fry,frz);
      const f2 visc_densi=mul2(mul2(pk1(FMUL(c.ddtkh,c.cs0)),sub2(mul2(pk1(rhop1),ovrhop2),pk1(1.f))),ovrr2eta);
      a.delta=add2(a.delta,mul2(mul2(visc_densi,dot3),mass2));
    }
  }
  if(TDENS==B2S_DDT_FOURTAKAS || (TDENS==B2S_DDT_FOURTAKASFULL && !BOUNDP2)){
    if(BOUNDP2)a.mark=true;
    else{
      const f2 dot3=dot3_2(drx,dry,drz,frx,fry,frz);
      const f2 one=pk1(1.f);
      const f2 x=sub2(fma2(pk1(c.ddtgz),drz,one),one);
      const f2 pw=fma2(x,fma2(x,fma2(x,pk1(c.ddt_c3),pk1(c.ddt_c2)),pk1(c.ddt_c1)),one);
      const f2 rho0=pk1(c.rhopzero);
      const f2 drho=sub2(mul2(rho0,pw),rho0);
      const f2 visc_densi=mul2(mul2(pk1(FMUL(c.ddtkh,c.cs0)),sub2(pk(FSUB(va.w,rhop1),FSUB(vb.w,rhop1)),drho)),ovrr2eta);
      a.delta=sub2(a.delta,mul2(mul2(mul2(visc_densi,dot3),mass2),ovrhop2));
    }
  }
  //-Viscosity (artificial).
  {
    const f2 dot=dot3_2(drx,dry,drz,dvx,dvy,dvz);
    const f2 dot_rr2=mul2(dot,ovrr2eta);
    const f2 dm=mul2(dot_rr2,m2);                       // misses contribute 0 to the maximum (visc starts at 0)
    a.visc=fmaxf(a.visc,fmaxf(lo(dm),hi(dm)));
    const f2 amubar=mul2(pk1(c.kernelh),dot_rr2);
    const f2 robar=mul2(pk(FADD(rhop1,va.w),FADD(rhop1,vb.w)),pk1(0.5f));
    const f2 pi_visc=mul2(mul2(mul2(pk1(FMUL(-visco,c.cs0)),amubar),pk(b2s_rcp(lo(robar)),b2s_rcp(hi(robar)))),mass2);
    const f2 npi=pk(lo(dot)<0.f? -lo(pi_visc): 0.f, hi(dot)<0.f? -hi(pi_visc): 0.f);   // only approaching pairs (dot<0)
    a.ax=fma2(npi,frx,a.ax); a.ay=fma2(npi,fry,a.ay); a.az=fma2(npi,frz,a.az);
  }
}

// ROWMODE 0: one neighbour row at a time.  ROWMODE 1 (variant 5): every row together with its point mirror (y,z)<->(-y,-z): a lane
// that sits close to one side of its cell has many hits in one of the two rows and few in the other, so the sum is much flatter
// across the lanes of a warp than either count and fewer rounds are lost to the lane with most hits (tools/analyze_rounds.py:
// 121 -> 103 rounds per particle).  Hits are still paired inside a row, so a round needs one set of row constants.
template<int TDENS,int ROWMODE,int MINB>
__global__ void __launch_bounds__(P2_BS,MINB) k_forces_fluid_pair2(ForcesArgs g, const __grid_constant__ DevCte c)
{
  const DivState* ds=g.ds;
  const uint32_t npb=ds->npb, np=ds->np;
  const uint32_t p1=npb+blockIdx.x*blockDim.x+threadIdx.x;
  // slab halo ghosts (CODE_PERIODIC) are neighbours only: their own sums would be thrown away by the integrator
  const bool valid=(p1<np && B2S_CODE_ISNORMAL(g.code[p1]));
  const uint32_t plast=np-1u;
  float ace2=0.f, viscmax=0.f;
  float4 pp1=make_float4(0,0,0,0), velrhop1=make_float4(0,0,0,1.f);
  CellRange r; r.cx=r.cy=r.cz=0; r.xini=r.xfin=r.yini=r.yfin=r.zini=r.zfin=0;
  if(valid){ pp1=g.posp[p1]; velrhop1=g.velrho[p1]; r=cell_range(g.dcell[p1],c,ds); }
  const float pressp1=pp1.w;
  const int d=c.scelldiv, ncx=int(ds->ncx), nsheet=int(ds->nsheet), cmx=int(ds->cellmin[0]);
  const float xin=pp1.x-c.scell*float((r.cx+cmx)&((1<<B2S_XBLK)-1));   // x inside the own cell (row_trim only)
  // the untrimmed x window touches at most two blocks of 2^B2S_XBLK cells: A (of xini) and B (of xfin-1)
  const bool oneblk=(((r.xini+cmx)>>B2S_XBLK)==((r.xfin-1+cmx)>>B2S_XBLK));
  const int xmid0=(((r.xfin-1+cmx)>>B2S_XBLK)<<B2S_XBLK)-cmx;          // first cell of block B
  const float sxA=FADD(pp1.x,b2s_offx(c,r.cx+cmx,r.xini+cmx)), sxB=FADD(pp1.x,b2s_offx(c,r.cx+cmx,r.xfin-1+cmx));
  // FP16 pre-filter constants (cell units): target x for the two blocks, threshold with margin
  const float ovs=1.0f/c.scell;
  const uint32_t jlast=(np-1u)>>1;
  const __half hA=__float2half_rn(sxA*ovs), hB=__float2half_rn(sxB*ovs);
  const uint32_t hAA=h2u(__halves2half2(hA,hA)), hBB=h2u(__halves2half2(hB,hB)), hAB=h2u(__halves2half2(hA,hB));
  const float kc=c.kernelsize2*ovs*ovs;
  const uint32_t thr2=h2u(__float2half2_rn(kc*1.015f+0.02f));
  // The boundary pass runs FIRST and its six sums are parked in shared memory, so that only one set of packed accumulators
  // is alive in the (long) fluid pass: fewer registers, more resident warps.  finalize_store gets both sets as before.
  __shared__ float stash[6][P2_BS];
  Acc2 acc;
  acc.ax=acc.ay=acc.az=acc.ar=acc.delta=0ull; acc.visc=0.f; acc.mark=false;
  // Whole boundary pass skipped for warps whose particles have no boundary particle in their neighbour windows (one byte per
  // particle from the per-cell flags of the divide instead of 2 x 25 begin-table loads): most of the fluid is far from a wall.
  const bool nearb=(g.nearbound==nullptr || (valid && g.nearbound[r.cx+r.cy*ncx+r.cz*nsheet]!=0));
  const bool warpnearb=(__any_sync(0xffffffffu,nearb)!=0);
  #pragma unroll 1
  for(int pass=1;pass>=0;pass--){
    if(pass==1 && !warpnearb)continue;   // (the stash below is only read when it was written: see the store after the loop)
    const uint32_t* beg=(pass==0? g.beg+ds->box_fluid: g.beg);   // pass 1: boundary neighbours (first), pass 0: fluid neighbours
    // Rows (z,y) of the neighbour window, software-pipelined: the begin-table loads of row i+1 are issued before row i is
    // processed, so their latency (three dependent loads after the trim) hides behind the pair work of the current row.
    struct Row { uint32_t pbeg,pend,pmid; float sy,sz; };
    auto setup_row=[&](int z,int y)->Row{
      Row w; w.pbeg=w.pend=w.pmid=0;
      w.sz=FADD(pp1.z,c.scell*float(r.cz-z)); w.sy=FADD(pp1.y,c.scell*float(r.cy-y));
      int xa,xb;
      if(z>=r.zini && z<r.zfin && y>=r.yini && y<r.yfin){
        const uint32_t* brow=beg+(y*ncx+z*nsheet);
        // boundary pass: most fluid particles have no boundary particle anywhere in the row -> two loads settle it, no trim
        if((pass==0 || brow[r.xini]!=brow[r.xfin]) && row_trim(c,r,xin,w.sy,w.sz,xa,xb)){
          w.pbeg=brow[xa]; w.pend=brow[xb];
          w.pmid=(oneblk || xmid0<=xa? w.pbeg: (xmid0>=xb? w.pend: brow[xmid0]));
        }
      }
      return w;
    };
    const int nrows=(2*d+1)*(2*d+1);
#if P2_MASK==2 && P2_PIPE==0
    if(ROWMODE==1){
      //-FP16 pre-filter of one chunk of a row (same test as in the single-row loop below).
#if P2_CHK2
      auto check=[&](const Row& w,uint32_t jb,uint32_t j0,uint32_t n)->uint32_t{
        const uint32_t pb0=(jb+j0)<<1;
        if(pb0>=w.pend)return 0u;                          // this lane's row ended before the chunk (empty rows: pbeg=pend=0)
        const uint32_t sy2=h2u(__float2half2_rn(w.sy*ovs)), sz2=h2u(__float2half2_rn(w.sz*ovs));
        // No clamp of the pair index: the lane has a particle in the chunk, so the 16 pairs end at most 15 pairs behind the last
        // one of the array, inside the padding of posh (ctx.cu); whatever is read there cannot set a bit that survives the range
        // mask below.  The bits are immediates (full unroll), and x-block B starts at particle nA of the chunk: a pair is taken as
        // (A,A) or (B,B) in the loop, the one pair that straddles (nA odd) is redone with (A,B) afterwards.
        const uint4* __restrict__ q0=g.posh+(jb+j0);
        const int nA=int(w.pmid)-int(pb0);
        uint32_t mask=0;
        auto one=[&](uint32_t k,uint32_t sx2){
          const uint4 q=__ldg(q0+k);
          const __half2 dx=__hsub2(u2h(sx2),u2h(q.x)), dy=__hsub2(u2h(sy2),u2h(q.y)), dz=__hsub2(u2h(sz2),u2h(q.z));
          const __half2 rr=__hfma2(dz,dz,__hfma2(dy,dy,__hmul2(dx,dx)));
          hit2_or(mask,h2u(rr),thr2,1u<<(2u*k),2u<<(2u*k));
        };
        #pragma unroll
        for(uint32_t k0=0;k0<16u;k0+=4u){
          if(k0<n){
            #pragma unroll
            for(uint32_t k=k0;k<k0+4u;k++)one(k,(int(2u*k+1u)<nA? hAA: hBB));
          }
        }
        if((nA&1) && nA>0 && nA<32){
          const uint32_t k=uint32_t(nA)>>1;
          mask&=~(3u<<(2u*k));
          one(k,hAB);
        }
        if(pb0<w.pbeg)mask&=~1u;
        { const uint32_t v=w.pend-pb0; if(v<32u)mask&=((1u<<v)-1u); }
        return mask;
      };
#else
      auto check=[&](const Row& w,uint32_t jb,uint32_t j0,uint32_t n)->uint32_t{
        const uint32_t pb0=(jb+j0)<<1;
        const uint32_t sy2=h2u(__float2half2_rn(w.sy*ovs)), sz2=h2u(__float2half2_rn(w.sz*ovs));
        uint32_t mask=0, bit=1u;
        #pragma unroll 4
        for(uint32_t k=0;k<n;k++,bit<<=2){
          const uint32_t p=pb0+2u*k;
          const uint4 q=__ldg(g.posh+min(jb+j0+k,jlast));
          const uint32_t sx2=(p+1u<w.pmid? hAA: (p>=w.pmid? hBB: hAB));
          const __half2 dx=__hsub2(u2h(sx2),u2h(q.x)), dy=__hsub2(u2h(sy2),u2h(q.y)), dz=__hsub2(u2h(sz2),u2h(q.z));
          const __half2 rr=__hfma2(dz,dz,__hfma2(dy,dy,__hmul2(dx,dx)));
          hit2_or(mask,h2u(rr),thr2,bit,bit+bit);
        }
        if(pb0<w.pbeg)mask&=~1u;
        { const uint32_t v=(w.pend>pb0? w.pend-pb0: 0u); if(v<32u)mask&=((1u<<v)-1u); }
        return mask;
      };
#endif
      const int half=nrows>>1;
      int za=r.cz-d, ya=r.cy-d, zb=r.cz+d, yb=r.cy+d;     // row i walks forward from the first row, its mirror backward from the last
      Row nextA=setup_row(za,ya);
      #pragma unroll 1
      for(int i=0;i<=half;i++){
        const Row A=nextA;
        Row B; B.pbeg=B.pend=B.pmid=0; B.sy=B.sz=0.f;
        if(i<half)B=setup_row(zb,yb);                      // the centre row (i==half) has no mirror
        if(++ya>r.cy+d){ ya=r.cy-d; za++; }
        if(--yb<r.cy-d){ yb=r.cy+d; zb--; }
        if(i<half)nextA=setup_row(za,ya);
        const uint32_t jbA=A.pbeg>>1, npA=__reduce_max_sync(0xffffffffu,((A.pend+1u)>>1)-jbA);
        const uint32_t jbB=B.pbeg>>1, npB=__reduce_max_sync(0xffffffffu,((B.pend+1u)>>1)-jbB);
        const uint32_t npmax=max(npA,npB);
        #pragma unroll 1
        for(uint32_t j0=0;j0<npmax;j0+=16){
          uint32_t maskA=(j0<npA? check(A,jbA,j0,min(16u,npA-j0)): 0u);
          uint32_t maskB=(j0<npB? check(B,jbB,j0,min(16u,npB-j0)): 0u);
          const uint32_t mine=((uint32_t(__popc(maskA))+1u)>>1)+((uint32_t(__popc(maskB))+1u)>>1);
          #pragma unroll 1
          for(uint32_t it=__reduce_max_sync(0xffffffffu,mine);it;it--){
            const bool ua=(maskA!=0u);
            uint32_t m=(ua? maskA: maskB);
            if(m){
              const uint32_t ka=__ffs(m)-1u; m&=m-1u;
              const bool two=(m!=0u);
              const uint32_t kb=(two? __ffs(m)-1u: ka); m&=m-1u;
              if(ua)maskA=m; else maskB=m;
              const uint32_t pb0=((ua? jbA: jbB)+j0)<<1, pmid=(ua? A.pmid: B.pmid);
              const float sy=(ua? A.sy: B.sy), sz=(ua? A.sz: B.sz);
              const uint32_t p2a=pb0+ka, p2b=pb0+kb;
              const float4 pa=__ldg(g.posp+p2a), pb=__ldg(g.posp+p2b);
              const float4 va=__ldg(g.velrho+p2a), vb=__ldg(g.velrho+p2b);
              const f2 drx=pk(FSUB(p2a>=pmid? sxB: sxA,pa.x),FSUB(p2b>=pmid? sxB: sxA,pb.x));
              const f2 dry=pk(FSUB(sy,pa.y),FSUB(sy,pb.y)), drz=pk(FSUB(sz,pa.z),FSUB(sz,pb.z));
              const f2 rr2=fma2(drz,drz,fma2(dry,dry,mul2(drx,drx)));   // b2s_rr2 order
              const float ra=lo(rr2), rb=hi(rr2);
              const bool ha=(ra<=c.kernelsize2 && ra>=B2S_ALMOSTZERO), hb=(two && rb<=c.kernelsize2 && rb>=B2S_ALMOSTZERO);
#if P2_VNOW
              if((ha||hb) && va.w==va.w && vb.w==vb.w){   // the branch waits for the velrho gathers too: ptxas cannot sink them below it
#else
              if(ha||hb){
#endif
                if(pass==0)pair2_body<TDENS,false>(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,va,vb,pa.w,pb.w,c.visco_f,acc);
                else       pair2_body<TDENS,true >(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,va,vb,pa.w,pb.w,c.visco_b,acc);
              }
            }
          }
        }
      }
    }
    else
#endif
    {
    int zn=r.cz-d, yn=r.cy-d;
    Row nextrow=setup_row(zn,yn);
    #pragma unroll 1
    for(int irow=0;irow<nrows;irow++){
      {
        const Row w=nextrow;
        if(++yn>r.cy+d){ yn=r.cy-d; zn++; }
        if(irow+1<nrows)nextrow=setup_row(zn,yn);
        const uint32_t pbeg=w.pbeg, pend=w.pend, pmid=w.pmid;
        const float sy=w.sy, sz=w.sz;
        const uint32_t lenmax=__reduce_max_sync(0xffffffffu,pend-pbeg);
#if P2_MASK==2
        //-Two phases per chunk of 16 candidate PAIRS (32 particles, aligned to even indices):
        // (1) FP16 pre-filter in cell units on the pair-interleaved posh array: one 16-byte load and ~8 instructions per
        //     candidate pair; the threshold carries a margin above the FP16 error (b2s_internal.h posh), so no real pair is lost;
        // (2) the set bits are evaluated two at a time in FP32 exactly like variant 0 (false positives drop out at the exact test).
        const uint32_t jb=pbeg>>1, npair=((pend+1u)>>1)-jb;
        const uint32_t npmax=__reduce_max_sync(0xffffffffu,npair);
        const uint32_t sy2=h2u(__float2half2_rn(sy*ovs)), sz2=h2u(__float2half2_rn(sz*ovs));
        #pragma unroll 1
        for(uint32_t j0=0;j0<npmax;j0+=16){
          const uint32_t n=min(16u,npmax-j0), pb0=(jb+j0)<<1;
          uint32_t mask=0, bit=1u;
          #pragma unroll 4
          for(uint32_t k=0;k<n;k++,bit<<=2){
            const uint32_t p=pb0+2u*k;
            const uint4 w=__ldg(g.posh+min(jb+j0+k,jlast));
            const uint32_t sx2=(p+1u<pmid? hAA: (p>=pmid? hBB: hAB));
            const __half2 dx=__hsub2(u2h(sx2),u2h(w.x)), dy=__hsub2(u2h(sy2),u2h(w.y)), dz=__hsub2(u2h(sz2),u2h(w.z));
            const __half2 rr=__hfma2(dz,dz,__hfma2(dy,dy,__hmul2(dx,dx)));
            hit2_or(mask,h2u(rr),thr2,bit,bit+bit);
          }
          //-Bits of particles outside [pbeg,pend) (first/last pair of the row) are cleared.
          if(pb0<pbeg)mask&=~1u;
          { const uint32_t v=(pend>pb0? pend-pb0: 0u); if(v<32u)mask&=((1u<<v)-1u); }
#if P2_PIPE==2
          //-Rounds with the position gathers one round ahead: the hit indices and posp of round i+1 are fetched before round i
          // is evaluated (8 registers of look-ahead instead of 16: velrho is loaded inside the round, behind the dr/rr2 work).
          struct Hit { uint32_t p2a,p2b; bool ok,two; float4 pa,pb; };
          auto fetch=[&](Hit& h){
            h.ok=(mask!=0u);
            if(h.ok){
              const uint32_t ka=__ffs(mask)-1u; mask&=mask-1u;
              h.two=(mask!=0u);
              const uint32_t kb=(h.two? __ffs(mask)-1u: ka); mask&=mask-1u;
              h.p2a=pb0+ka; h.p2b=pb0+kb;
              h.pa=__ldg(g.posp+h.p2a); h.pb=__ldg(g.posp+h.p2b);
            }
          };
          auto eval=[&](const Hit& h){
            if(h.ok){
              const float4 va=__ldg(g.velrho+h.p2a), vb=__ldg(g.velrho+h.p2b);
              const f2 drx=pk(FSUB(h.p2a>=pmid? sxB: sxA,h.pa.x),FSUB(h.p2b>=pmid? sxB: sxA,h.pb.x));
              const f2 dry=pk(FSUB(sy,h.pa.y),FSUB(sy,h.pb.y)), drz=pk(FSUB(sz,h.pa.z),FSUB(sz,h.pb.z));
              const f2 rr2=fma2(drz,drz,fma2(dry,dry,mul2(drx,drx)));   // b2s_rr2 order
              const float ra=lo(rr2), rb=hi(rr2);
              const bool ha=(ra<=c.kernelsize2 && ra>=B2S_ALMOSTZERO), hb=(h.two && rb<=c.kernelsize2 && rb>=B2S_ALMOSTZERO);
              if(ha||hb){
                if(pass==0)pair2_body<TDENS,false>(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,va,vb,h.pa.w,h.pb.w,c.visco_f,acc);
                else       pair2_body<TDENS,true >(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,va,vb,h.pa.w,h.pb.w,c.visco_b,acc);
              }
            }
          };
          const uint32_t rounds=(__reduce_max_sync(0xffffffffu,uint32_t(__popc(mask)))+1u)>>1;
          Hit A,B;
          if(rounds)fetch(A);
          #pragma unroll 1
          for(uint32_t it=0;it<rounds;it+=2){
            fetch(B); eval(A);
            if(it+2<rounds)fetch(A);
            eval(B);
          }
        }
#elif P2_PIPE==1
          //-Rounds, software-pipelined two deep: the gathers (posp, velrho of both hits) of round i+1 are in flight while
          // round i is evaluated.  Two register sets A/B alternate (loop unrolled by two), no register copies.
          struct Hit { uint32_t p2a,p2b; bool ok,two; float4 pa,pb,va,vb; };
          auto fetch=[&](Hit& h){
            h.ok=(mask!=0u);
            if(h.ok){
              const uint32_t ka=__ffs(mask)-1u; mask&=mask-1u;
              h.two=(mask!=0u);
              const uint32_t kb=(h.two? __ffs(mask)-1u: ka); mask&=mask-1u;
              h.p2a=pb0+ka; h.p2b=pb0+kb;
              h.pa=__ldg(g.posp+h.p2a); h.pb=__ldg(g.posp+h.p2b); h.va=__ldg(g.velrho+h.p2a); h.vb=__ldg(g.velrho+h.p2b);
            }
          };
          auto eval=[&](const Hit& h){
            if(h.ok){
              const f2 drx=pk(FSUB(h.p2a>=pmid? sxB: sxA,h.pa.x),FSUB(h.p2b>=pmid? sxB: sxA,h.pb.x));
              const f2 dry=pk(FSUB(sy,h.pa.y),FSUB(sy,h.pb.y)), drz=pk(FSUB(sz,h.pa.z),FSUB(sz,h.pb.z));
              const f2 rr2=fma2(drz,drz,fma2(dry,dry,mul2(drx,drx)));   // b2s_rr2 order
              const float ra=lo(rr2), rb=hi(rr2);
              const bool ha=(ra<=c.kernelsize2 && ra>=B2S_ALMOSTZERO), hb=(h.two && rb<=c.kernelsize2 && rb>=B2S_ALMOSTZERO);
              if(ha||hb){
                if(pass==0)pair2_body<TDENS,false>(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,h.va,h.vb,h.pa.w,h.pb.w,c.visco_f,acc);
                else       pair2_body<TDENS,true >(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,h.va,h.vb,h.pa.w,h.pb.w,c.visco_b,acc);
              }
            }
          };
          const uint32_t rounds=(__reduce_max_sync(0xffffffffu,uint32_t(__popc(mask)))+1u)>>1;
          Hit A,B;
          if(rounds)fetch(A);
          #pragma unroll 1
          for(uint32_t it=0;it<rounds;it+=2){
            fetch(B); eval(A);
            if(it+2<rounds)fetch(A);
            eval(B);
          }
        }
#else
          #pragma unroll 1
          for(uint32_t it=(__reduce_max_sync(0xffffffffu,uint32_t(__popc(mask)))+1u)>>1;it;it--){
            if(mask){
              const uint32_t ka=__ffs(mask)-1u; mask&=mask-1u;
              const bool two=(mask!=0u);
              const uint32_t kb=(two? __ffs(mask)-1u: ka); mask&=mask-1u;
              const uint32_t p2a=pb0+ka, p2b=pb0+kb;
              const float4 pa=__ldg(g.posp+p2a), pb=__ldg(g.posp+p2b);
              const float4 va=__ldg(g.velrho+p2a), vb=__ldg(g.velrho+p2b);   // issued with posp: one load latency per round, not two
              const f2 drx=pk(FSUB(p2a>=pmid? sxB: sxA,pa.x),FSUB(p2b>=pmid? sxB: sxA,pb.x));
              const f2 dry=pk(FSUB(sy,pa.y),FSUB(sy,pb.y)), drz=pk(FSUB(sz,pa.z),FSUB(sz,pb.z));
              const f2 rr2=fma2(drz,drz,fma2(dry,dry,mul2(drx,drx)));   // b2s_rr2 order
              const float ra=lo(rr2), rb=hi(rr2);
              const bool ha=(ra<=c.kernelsize2 && ra>=B2S_ALMOSTZERO), hb=(two && rb<=c.kernelsize2 && rb>=B2S_ALMOSTZERO);
              if(ha||hb){
                if(pass==0)pair2_body<TDENS,false>(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,va,vb,pa.w,pb.w,c.visco_f,acc);
                else       pair2_body<TDENS,true >(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,va,vb,pa.w,pb.w,c.visco_b,acc);
              }
            }
          }
        }
#endif
#elif P2_MASK==1
        //-Two phases per chunk of up to 32 candidates: (1) check every candidate and record the hits in a per-lane bit mask,
        // (2) evaluate the hits two at a time.  Phase 2 only iterates as long as the lane with most hits needs, so the packed
        // body runs on real pairs (both halves used) instead of on whatever two consecutive candidates happen to be.
        #pragma unroll 1
        for(uint32_t k0=0;k0<lenmax;k0+=32){
          const uint32_t n=min(32u,lenmax-k0), pb0=pbeg+k0;
          uint32_t mask=0;
          #pragma unroll 4
          for(uint32_t k=0;k<n;k++){
            const uint32_t p2=pb0+k;
            const float4 pp2=__ldg(g.posp+min(p2,plast));
            const float drx=FSUB(p2>=pmid? sxB: sxA,pp2.x), dry=FSUB(sy,pp2.y), drz=FSUB(sz,pp2.z);
            const float rr2=b2s_rr2(drx,dry,drz);
            if(p2<pend && rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO)mask|=(1u<<k);
          }
          #pragma unroll 1
          for(uint32_t it=(__reduce_max_sync(0xffffffffu,uint32_t(__popc(mask)))+1u)>>1;it;it--){
            if(mask){
              const uint32_t ka=__ffs(mask)-1u; mask&=mask-1u;
              const bool hb=(mask!=0u);
              const uint32_t kb=(hb? __ffs(mask)-1u: ka); mask&=mask-1u;
              const uint32_t p2a=pb0+ka, p2b=pb0+kb;
              const float4 pa=__ldg(g.posp+p2a), pb=__ldg(g.posp+p2b);
              const float4 va=__ldg(g.velrho+p2a), vb=__ldg(g.velrho+p2b);
              const f2 drx=pk(FSUB(p2a>=pmid? sxB: sxA,pa.x),FSUB(p2b>=pmid? sxB: sxA,pb.x));
              const f2 dry=pk(FSUB(sy,pa.y),FSUB(sy,pb.y)), drz=pk(FSUB(sz,pa.z),FSUB(sz,pb.z));
              const f2 rr2=fma2(drz,drz,fma2(dry,dry,mul2(drx,drx)));   // b2s_rr2 order
              if(pass==0)pair2_body<TDENS,false>(c,drx,dry,drz,rr2,true,hb,velrhop1,pressp1,va,vb,pa.w,pb.w,c.visco_f,acc);
              else       pair2_body<TDENS,true >(c,drx,dry,drz,rr2,true,hb,velrhop1,pressp1,va,vb,pa.w,pb.w,c.visco_b,acc);
            }
          }
        }
#else
        #pragma unroll 2
        for(uint32_t k=0;k<lenmax;k+=2){
          const uint32_t p2a=pbeg+k, p2b=p2a+1u;
          const uint32_t qa=min(p2a,plast), qb=min(p2b,plast);      // loads always hit real particles (finite data)
          const float4 pa=__ldg(g.posp+qa), pb=__ldg(g.posp+qb);
#if P2_EARLYV
          const float4 va=__ldg(g.velrho+qa), vb=__ldg(g.velrho+qb);
#endif
          const f2 drx=pk(FSUB(p2a>=pmid? sxB: sxA,pa.x),FSUB(p2b>=pmid? sxB: sxA,pb.x));
          const f2 dry=pk(FSUB(sy,pa.y),FSUB(sy,pb.y)), drz=pk(FSUB(sz,pa.z),FSUB(sz,pb.z));
          const f2 rr2=fma2(drz,drz,fma2(dry,dry,mul2(drx,drx)));   // b2s_rr2 order
          const float ra=lo(rr2), rb=hi(rr2);
          const bool ha=(p2a<pend && ra<=c.kernelsize2 && ra>=B2S_ALMOSTZERO);
          const bool hb=(p2b<pend && rb<=c.kernelsize2 && rb>=B2S_ALMOSTZERO);
          if(ha||hb){
#if !P2_EARLYV
            const float4 va=__ldg(g.velrho+qa), vb=__ldg(g.velrho+qb);
#endif
            if(pass==0)pair2_body<TDENS,false>(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,va,vb,pa.w,pb.w,c.visco_f,acc);
            else       pair2_body<TDENS,true >(c,drx,dry,drz,rr2,ha,hb,velrhop1,pressp1,va,vb,pa.w,pb.w,c.visco_b,acc);
          }
        }
#endif
      }
    }
    }
    if(pass==1){ // park the boundary sums, restart the accumulators for the fluid pass
      stash[0][threadIdx.x]=FADD(lo(acc.ax),hi(acc.ax)); stash[1][threadIdx.x]=FADD(lo(acc.ay),hi(acc.ay)); stash[2][threadIdx.x]=FADD(lo(acc.az),hi(acc.az));
      stash[3][threadIdx.x]=FADD(lo(acc.ar),hi(acc.ar)); stash[4][threadIdx.x]=(acc.mark? FLT_MAX: FADD(lo(acc.delta),hi(acc.delta))); stash[5][threadIdx.x]=acc.visc;
      acc.ax=acc.ay=acc.az=acc.ar=acc.delta=0ull; acc.visc=0.f; acc.mark=false;
    }
  }
  if(valid){
    PairAcc sf={0,0,0,0,0,0,0,0,0,0,0,0}, sb={0,0,0,0,0,0,0,0,0,0,0,0};
    sf.ax=FADD(lo(acc.ax),hi(acc.ax)); sf.ay=FADD(lo(acc.ay),hi(acc.ay)); sf.az=FADD(lo(acc.az),hi(acc.az)); sf.ar=FADD(lo(acc.ar),hi(acc.ar));
    sf.delta=(acc.mark? FLT_MAX: FADD(lo(acc.delta),hi(acc.delta))); sf.visc=acc.visc;
    if(warpnearb){
      sb.ax=stash[0][threadIdx.x]; sb.ay=stash[1][threadIdx.x]; sb.az=stash[2][threadIdx.x]; sb.ar=stash[3][threadIdx.x];
      sb.delta=stash[4][threadIdx.x]; sb.visc=stash[5][threadIdx.x];
    }
    finalize_store<B2S_VISCO_ARTIFICIAL,TDENS>(g,c,p1,velrhop1,sf,sb,ace2,viscmax);
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
bool b2s_forces_pair2_supported(const DevCte& c){
  return c.tkernel==B2S_KERNEL_WENDLAND && c.tvisco==B2S_VISCO_ARTIFICIAL && !c.use_floating && !c.mdbc2 && (2*c.scelldiv+1)<=5;
}

int b2s_forces_pair2_launch(b2s_ctx* ctx, const ForcesArgs& g){
  const DevCte& c=ctx->cte;
  const uint32_t npf=ctx->np_host-ctx->npb_host;
  const unsigned gb=b2s_blocks(npf,P2_BS);
  #define P2_LAUNCH(RM,MB) switch(c.tdensity){ \
    case B2S_DDT_NONE:      k_forces_fluid_pair2<B2S_DDT_NONE,RM,MB><<<gb,P2_BS,0,ctx->stream>>>(g,c); break; \
    case B2S_DDT_MOLTENI:   k_forces_fluid_pair2<B2S_DDT_MOLTENI,RM,MB><<<gb,P2_BS,0,ctx->stream>>>(g,c); break; \
    case B2S_DDT_FOURTAKAS: k_forces_fluid_pair2<B2S_DDT_FOURTAKAS,RM,MB><<<gb,P2_BS,0,ctx->stream>>>(g,c); break; \
    default:                k_forces_fluid_pair2<B2S_DDT_FOURTAKASFULL,RM,MB><<<gb,P2_BS,0,ctx->stream>>>(g,c); break; }
  // variant 5 = mirror-row pairing.  Builds for 6 and 8 resident blocks per SM were timed too (80 registers: 11.75 ms, 64 registers
  // with 400 bytes of spills: 11.46 ms, against 11.42 ms at 7 blocks / 72 registers; profiles/r01_pair2_tuning.txt)
  if(ctx->variant==5){ P2_LAUNCH(1,P2_MINB) }
  else{ P2_LAUNCH(0,P2_MINB) }
  #undef P2_LAUNCH
  B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
