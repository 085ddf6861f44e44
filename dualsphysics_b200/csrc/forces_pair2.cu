// forces_pair2.cu -- interaction variant 5: one thread per particle, FP16 pre-filter into per-lane hit masks, two hits per lane
// and round in packed FP32 (sm_100a f32x2), every neighbour row walked together with its point mirror.
//
// Same physics as variant 0 (forces.cu / forces_common.cuh pair_fluid), which restates cusph::Interaction_Forces /
// KerInteractionForcesFluid (src/source/JSphGpu_ker.cu:673-964) and JSphCpu::InteractionForcesFluid (src/source/JSphCpu.cpp:736-1125).
//
// Why: variant 0 is bound by FP32 *issue slots* (profiles/README.md: 80% of the slots busy, FMA pipe 45%, 19 of 32 lanes active
// because the ~70-instruction pair body sits under the rr2<(2h)^2 branch).  Blackwell's fma/mul/add.rn.f32x2 do two FP32
// operations per issue slot (profiles/r02b_microbench.txt: same FMA throughput as FFMA, half the slots).  Per chunk of a row a
// lane first collects its hits in a bit mask (FP16 test on the pair-interleaved posh copy, margin above the FP16 error), then
// evaluates the set bits two at a time in f32x2 registers with pack_body (forces_pair2.cuh), which needs no exact distance test:
// beyond 2h the kernel factor clamps to zero.
//   * rows of neighbour cells are walked as ONE index range (block-relative x, b2s_internal.h B2S_XBLK) and trimmed to the
//     cells the kernel sphere reaches (row_trim);
//   * row i of the window is processed together with row (last-i), its point mirror (y,z)<->(-y,-z): a lane close to one side
//     of its cell has many hits in one of the two and few in the other, so the sum is flat across the warp and fewer rounds are
//     lost to the lane with most hits (tools/analyze_rounds.py: 121 -> 103 rounds per particle);
//   * ROW DESCRIPTORS: the window is walked in phases of one z plane and its mirror plane.  At the head of a phase a lane works
//     out the rows of all its turns (trim, then three begin-table loads per row that depend on nothing but the trim) and parks them
//     in its own shared-memory slots; a turn then costs one LDS.128.  The former set-up of every row right before its use (two
//     set-ups of ~130 instructions per turn, each with three dependent loads and the reloads of the spilled cell range) cost
//     more in stalls than in instructions: issue slots 67 -> 74% busy, 72 -> 64 registers (8 blocks per SM), 9.07 -> 7.95 ms
//     at 10M particles (profiles/README.md, r02E-r02K);
//   * every axis of the reference's dispatch is a template parameter (kernel, viscosity, DDT, floating bodies, mDBC2).
// History of the rejected builds (consecutive candidates instead of masks, software-pipelined rounds, velrho with the positions,
// single rows = the former variant 4, rows set up one ahead of their use): profiles/r01_pair2_tuning.txt, profiles/r02_pair2_ab.txt,
// profiles/README.md.
#include "forces_pair2.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

#ifndef P2_BS
#define P2_BS 128
#endif
#ifndef P2_MINB_SPS
#define P2_MINB_SPS 5
#endif
#ifndef P2_MINB_OTHER
#define P2_MINB_OTHER 8
#endif
#ifndef P2_MINB
#define P2_MINB 8            // resident blocks per SM the register allocation aims at (64 registers)
#endif

// FP16 pre-filter: rr2 of two candidates (half2) against the threshold; sets bit0/bit1 in mask for lo/hi
__device__ __forceinline__ void hit2_or(uint32_t& mask,uint32_t rr2,uint32_t thr2,uint32_t bit0,uint32_t bit1){
  asm("{\n .reg .pred p,q;\n setp.le.f16x2 p|q, %1, %2;\n @p or.b32 %0, %0, %3;\n @q or.b32 %0, %0, %4;\n}"
      :"+r"(mask):"r"(rr2),"r"(thr2),"r"(bit0),"r"(bit1));
}
__device__ __forceinline__ uint32_t h2u(__half2 h){ return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u){ return *reinterpret_cast<__half2*>(&u); }

// Resident blocks per SM the register allocation aims at (measured, profiles/README.md r02E-r02H): 8 (64 registers) for everything
// but laminar+SPS, which carries 6 more packed sums and 12 stress components per round: 5 (96 registers).
template<int TKER,int TVISCO,bool FLOATING> struct P2Cfg {
  static constexpr int MINB=(TVISCO==B2S_VISCO_LAMINARSPS? P2_MINB_SPS: ((FLOATING || TKER!=B2S_KERNEL_WENDLAND || TVISCO!=B2S_VISCO_ARTIFICIAL)? P2_MINB_OTHER: P2_MINB));
};

template<int TKER,int TVISCO,int TDENS,bool FLOATING,bool MDBC2>
__global__ void __launch_bounds__(P2_BS,(P2Cfg<TKER,TVISCO,FLOATING>::MINB)) k_forces_fluid_pair2(ForcesArgs g, const PackCte k, const __grid_constant__ DevCte c)
{
  __shared__ uint4 sdesc[5][P2_BS];   // row descriptors of one phase, one slot per thread and turn (see the phase loop)
  const DivState* ds=g.ds;
  const uint32_t npb=ds->npb, np=ds->np;
  const uint32_t p1=npb+blockIdx.x*blockDim.x+threadIdx.x;
  // slab halo ghosts (CODE_PERIODIC) are neighbours only: their own sums would be thrown away by the integrator
  const bool valid=(p1<np && B2S_CODE_ISNORMAL(g.code[p1]));
  float ace2=0.f, viscmax=0.f;
  float4 pp1=make_float4(0,0,0,0), velrhop1=make_float4(0,0,0,1.f);
  CellRange r; r.cx=r.cy=r.cz=0; r.xini=r.xfin=r.yini=r.yfin=r.zini=r.zfin=0;
  if(valid){ pp1=g.posp[p1]; velrhop1=g.velrho[p1]; r=cell_range(g.dcell[p1],c,ds); }
  const float pressp1=pp1.w, ovrhop1=b2s_rcp(velrhop1.w);
  PackTarget tg; tg.tensil=0.f; tg.ft=false;
  #pragma unroll
  for(int q=0;q<6;q++)tg.tau[q]=0.f;
  if(TKER==B2S_KERNEL_CUBIC)tg.tensil=FMUL(FMUL(pressp1,b2s_rcp(FMUL(velrhop1.w,velrhop1.w))),(pressp1>0.f? 0.01f: -0.2f));
  if(FLOATING && valid)tg.ft=B2S_CODE_ISFLOATING(g.code[p1]);
  if(TVISCO==B2S_VISCO_LAMINARSPS && valid){
    const float2* tp=(const float2*)(g.tau+size_t(p1)*6);
    const float2 t0=tp[0],t1=tp[1],t2=tp[2];
    tg.tau[0]=t0.x; tg.tau[1]=t0.y; tg.tau[2]=t1.x; tg.tau[3]=t1.y; tg.tau[4]=t2.x; tg.tau[5]=t2.y;
  }
  const int d=c.scelldiv, ncx=int(ds->ncx), nsheet=int(ds->nsheet), cmx=int(ds->cellmin[0]);
  const float xin=pp1.x-c.scell*float((r.cx+cmx)&((1<<B2S_XBLK)-1));   // x inside the own cell (row_trim only)
  // the untrimmed x window touches at most two blocks of 2^B2S_XBLK cells: A (of xini) and B (of xfin-1)
  const bool oneblk=(((r.xini+cmx)>>B2S_XBLK)==((r.xfin-1+cmx)>>B2S_XBLK));
  const int xmid0=(((r.xfin-1+cmx)>>B2S_XBLK)<<B2S_XBLK)-cmx;          // first cell of block B
  const float sxA=FADD(pp1.x,b2s_offx(c,r.cx+cmx,r.xini+cmx)), sxB=FADD(pp1.x,b2s_offx(c,r.cx+cmx,r.xfin-1+cmx));
  // FP16 pre-filter constants (cell units): target x for the two blocks, threshold with margin
  const float ovs=1.0f/c.scell;
  const __half hA=__float2half_rn(sxA*ovs), hB=__float2half_rn(sxB*ovs);
  const uint32_t hAA=h2u(__halves2half2(hA,hA)), hBB=h2u(__halves2half2(hB,hB)), hAB=h2u(__halves2half2(hA,hB));
  const float kc=c.kernelsize2*ovs*ovs;
  const uint32_t thr2=h2u(__float2half2_rn(kc*1.015f+0.02f));
  PackAcc acc;
  pack_zero(acc);
  uint32_t lenor=0;   // OR of the row lengths
  PairAcc sbnd={0,0,0,0,0,0,0,0,0,0,0,0};
  // Whole boundary pass skipped for warps whose particles have no boundary particle in their neighbour windows (one byte per
  // particle from the per-cell flags of the divide instead of 2 x 25 begin-table loads): most of the fluid is far from a wall.
  const bool nearb=(g.nearbound==nullptr || (valid && g.nearbound[r.cx+r.cy*ncx+r.cz*nsheet]!=0));
  const bool warpnearb=(__any_sync(0xffffffffu,nearb)!=0);
  #pragma unroll 1
  for(int pass=1;pass>=0;pass--){
    if(pass==1 && !warpnearb)continue;
    const uint32_t* beg=(pass==0? g.beg+ds->box_fluid: g.beg);   // pass 1: boundary neighbours (first), pass 0: fluid neighbours
    struct Row { uint32_t pbeg,pend,pmid; float sy,sz; };
    //-FP16 pre-filter of one chunk (16 candidate pairs) of a row.
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
      auto one=[&](uint32_t kk,uint32_t sx2){
        const uint4 q=__ldg(q0+kk);
        const __half2 dx=__hsub2(u2h(sx2),u2h(q.x)), dy=__hsub2(u2h(sy2),u2h(q.y)), dz=__hsub2(u2h(sz2),u2h(q.z));
        const __half2 rr=__hfma2(dz,dz,__hfma2(dy,dy,__hmul2(dx,dx)));
        hit2_or(mask,h2u(rr),thr2,1u<<(2u*kk),2u<<(2u*kk));
      };
      #pragma unroll
      for(uint32_t k0=0;k0<16u;k0+=4u){
        if(k0<n){
          #pragma unroll
          for(uint32_t kk=k0;kk<k0+4u;kk++)one(kk,(int(2u*kk+1u)<nA? hAA: hBB));
        }
      }
      if((nA&1) && nA>0 && nA<32){
        const uint32_t kk=uint32_t(nA)>>1;
        mask&=~(3u<<(2u*kk));
        one(kk,hAB);
      }
      if(pb0<w.pbeg)mask&=~1u;
      { const uint32_t v=w.pend-pb0; if(v<32u)mask&=((1u<<v)-1u); }
      return mask;
    };
    // Row descriptors.  The window is walked in phases of one z plane and its mirror plane: first a lane works out the rows of
    // ALL turns of the phase -- trim to the sphere, then three begin-table loads per row that depend on nothing but the trim, so
    // the loads of the (up to) 10 rows of a phase are in flight together -- and parks them in its own shared-memory slots
    // (sdesc[turn][thread] = {pbegA, lenA | midA<<16, pbegB, lenB | midB<<16}; no other thread reads them: no barrier).  The turn
    // loop then costs one LDS.128 instead of two row set-ups with three dependent loads each.
    const int n1=2*d+1;
    #pragma unroll 1
    for(int ph=0;ph<=d;ph++){
      const int nw=(ph<d? n1: d+1);                      // the last phase is the centre plane: its second half mirrors its first
      const int zA=r.cz-d+ph, zB=r.cz+d-ph;
      float kzo=c.scell*float(d-ph);
      asm volatile("":"+f"(kzo));   // opaque: kept (in a uniform register) instead of being rebuilt from the loop counters in every round: 8.08 -> 7.95 ms
      const float szA=FADD(pp1.z,kzo), szB=FADD(pp1.z,-kzo);
      const bool zokA=(valid && zA>=r.zini && zA<r.zfin), zokB=(valid && zB>=r.zini && zB<r.zfin);
      const float dzA=fmaxf(0.f,fmaxf(szA-c.scell,-szA)), dzB=fmaxf(0.f,fmaxf(szB-c.scell,-szB));
      const float kzA=c.kernelsize2*1.0002f-dzA*dzA, kzB=c.kernelsize2*1.0002f-dzB*dzB;
      auto desc_row=[&](bool zok,int z,int y,float sy,float kz)->uint2{
        // row_trim (forces_common.cuh) without branches: an empty or outside row reads three times beg[0] and gets len 0
        const float dy=fmaxf(0.f,fmaxf(sy-c.scell,-sy));
        const float xh2=kz-dy*dy;
        const float xh=xh2*b2s_rsqrt(xh2)+c.scell*1e-4f;
        int xa=max(r.xini,r.cx+int(floorf((xin-xh)*ovs)));
        int xb=min(r.xfin,r.cx+int(floorf((xin+xh)*ovs))+1);
        const bool ok=(zok && y>=r.yini && y<r.yfin && xh2>0.f && xa<xb);
        uint32_t off=uint32_t(y*ncx+z*nsheet);             // 32-bit unsigned table indices: one IMAD.WIDE per address
        if(!ok){ xa=xb=0; off=0u; }
        const int xm=(oneblk? xa: min(max(xmid0,xa),xb));
        const uint32_t pbeg=beg[off+uint32_t(xa)], pend=beg[off+uint32_t(xb)], pmid=beg[off+uint32_t(xm)];
        const uint32_t len=pend-pbeg;
        lenor|=len;                                        // tested once at the end of the kernel
        return make_uint2(len? pbeg: 0u,len|((pmid-pbeg)<<16));   // empty rows at 0: an odd begin would cost a chunk turn
      };
      #pragma unroll
      for(int w=0;w<5;w++){
        if(w<nw){
          const float syA=FADD(pp1.y,c.scell*float(d-w)), syB=FADD(pp1.y,c.scell*float(w-d));
          const uint2 a=desc_row(zokA,zA,r.cy-d+w,syA,kzA);
          const uint2 b=desc_row(zokB && (ph<d || w<d),zB,r.cy+d-w,syB,kzB);   // the centre row has no mirror
          sdesc[w][threadIdx.x]=make_uint4(a.x,a.y,b.x,b.y);
        }
      }
      #pragma unroll 1
      for(int w=0;w<nw;w++){
        const uint4 dq=sdesc[w][threadIdx.x];
        float kyA=c.scell*float(d-w);
        asm volatile("":"+f"(kyA));
        Row A,B;
        A.pbeg=dq.x; A.pend=dq.x+(dq.y&0xffffu); A.pmid=dq.x+(dq.y>>16); A.sy=FADD(pp1.y,kyA); A.sz=szA;
        B.pbeg=dq.z; B.pend=dq.z+(dq.w&0xffffu); B.pmid=dq.z+(dq.w>>16); B.sy=FADD(pp1.y,-kyA); B.sz=szB;
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
              const float sg=(ua? 1.f: -1.f), sy=FFMA(sg,kyA,pp1.y), sz=FFMA(sg,kzo,pp1.z);   // = A.sy/A.sz or B.sy/B.sz: the mirror row sits at the negated offset
              const uint32_t p2a=pb0+ka, p2b=pb0+kb;
              const float4 pa=__ldg(g.posp+p2a), pb=__ldg(g.posp+p2b);
              const float4 va=__ldg(g.velrho+p2a), vb=__ldg(g.velrho+p2b);
              const f2 drx=pk(FSUB(p2a>=pmid? sxB: sxA,pa.x),FSUB(p2b>=pmid? sxB: sxA,pb.x));
              const f2 dry=pk(FSUB(sy,pa.y),FSUB(sy,pb.y)), drz=pk(FSUB(sz,pa.z),FSUB(sz,pb.z));
              f2 rr2=fma2(drz,drz,fma2(dry,dry,mul2(drx,drx)));   // b2s_rr2 order
              if(!two)rr2=pk(lo(rr2),1e30f);                      // the repeated hit of a lone round: far outside the kernel, no contribution
              // No exact test and no branch: false positives of the pre-filter (rr2>(2h)^2) and the particle itself (rr2=0) contribute 0.
              PackNeigh nb; nb.mu=pk1(1.f); nb.plain=pk1(1.f); nb.heavy_a=nb.heavy_b=false; nb.taua=nb.taub=nullptr;
              if(pass==0){
                if(FLOATING){
                  const typecode ca=g.code[p2a], cb=g.code[p2b];
                  const bool fla=B2S_CODE_ISFLOATING(ca), flb=B2S_CODE_ISFLOATING(cb);
                  float ma=1.f, mb=1.f;
                  if(fla){ const float m=g.ftomassp[B2S_CODE_TYPEVALUE(ca)]; ma=FMUL(m,k.ovmassf); nb.heavy_a=(m<=c.massf*1.2f); }
                  if(flb){ const float m=g.ftomassp[B2S_CODE_TYPEVALUE(cb)]; mb=FMUL(m,k.ovmassf); nb.heavy_b=(m<=c.massf*1.2f); }
                  nb.mu=pk(ma,mb); nb.plain=pk(fla? 0.f: 1.f,flb? 0.f: 1.f);
                }
                if(TVISCO==B2S_VISCO_LAMINARSPS){ nb.taua=(const float2*)(g.tau+size_t(p2a)*6); nb.taub=(const float2*)(g.tau+size_t(p2b)*6); }
                pack_body<TKER,TVISCO,TDENS,FLOATING,MDBC2,false>(c,k,drx,dry,drz,rr2,velrhop1,pressp1,ovrhop1,tg,va,vb,pa.w,pb.w,nb,acc);
              }
              else{
                if(MDBC2){   // mDBC2: no mass without a submerged ghost node, tangential velocity for the viscous terms (JSphCpu.cpp:851-853,990-995)
                  nb.tva=__ldg(g.tangenvel+p2a); nb.tvb=__ldg(g.tangenvel+p2b);
                  nb.mu=pk(g.boundmode[p2a]==B2S_BMODE_MDBC2OFF? 0.f: 1.f,g.boundmode[p2b]==B2S_BMODE_MDBC2OFF? 0.f: 1.f);
                }
                pack_body<TKER,TVISCO,TDENS,FLOATING,MDBC2,true>(c,k,drx,dry,drz,rr2,velrhop1,pressp1,ovrhop1,tg,va,vb,pa.w,pb.w,nb,acc);
              }
            }
          }
        }
      }
    }
    if(pass==1){ // scale and keep the boundary sums, restart the accumulators for the fluid pass
      pack_finish<TKER,TVISCO,TDENS,true>(c,k,acc,velrhop1.w,(FLOATING && tg.ft),sbnd);
      pack_zero(acc);
    }
  }
  if(valid){
    PairAcc sf={0,0,0,0,0,0,0,0,0,0,0,0};
    pack_finish<TKER,TVISCO,TDENS,false>(c,k,acc,velrhop1.w,(FLOATING && tg.ft),sf);
    finalize_store<TVISCO,TDENS>(g,c,p1,velrhop1,sf,sbnd,ace2,viscmax);
  }
  if(lenor>0xffffu)g.st->error=B2S_ERR_UNSUPPORTED;   // a trimmed row of more than 65535 particles does not fit its descriptor: not a state to go on with
  //-Fused maxima: ViscDtMax and AceMax^2 (replaces ReduMaxFloat x2 + cudaMemcpy, JSphGpu_ker.cu:75-125).
  reduce_maxima(g.st,viscmax,ace2);
}

// ------------------------------------------------------------------------------------------------
// Every axis of the reference's dispatch (JSphGpu_ker.cu:1080-1179) has a packed instantiation; only windows wider than 5x5 rows
// (cells smaller than h) are left to variant 0.
bool b2s_forces_pair2_supported(const DevCte& c){ return (2*c.scelldiv+1)<=5; }

typedef void (*pair2_kernel_t)(ForcesArgs, const PackCte, const DevCte);
template<int TKER,int TVISCO,int TDENS> static pair2_kernel_t p2_pick_ft(bool ft,bool m2){
  if(m2)return ft? (pair2_kernel_t)k_forces_fluid_pair2<TKER,TVISCO,TDENS,true,true>: (pair2_kernel_t)k_forces_fluid_pair2<TKER,TVISCO,TDENS,false,true>;
  return ft? (pair2_kernel_t)k_forces_fluid_pair2<TKER,TVISCO,TDENS,true,false>: (pair2_kernel_t)k_forces_fluid_pair2<TKER,TVISCO,TDENS,false,false>;
}
template<int TKER,int TVISCO> static pair2_kernel_t p2_pick_dens(int td,bool ft,bool m2){
  switch(td){
    case B2S_DDT_NONE:      return p2_pick_ft<TKER,TVISCO,B2S_DDT_NONE>(ft,m2);
    case B2S_DDT_MOLTENI:   return p2_pick_ft<TKER,TVISCO,B2S_DDT_MOLTENI>(ft,m2);
    case B2S_DDT_FOURTAKAS: return p2_pick_ft<TKER,TVISCO,B2S_DDT_FOURTAKAS>(ft,m2);
    default:                return p2_pick_ft<TKER,TVISCO,B2S_DDT_FOURTAKASFULL>(ft,m2);
  }
}
template<int TKER> static pair2_kernel_t p2_pick_visco(int tv,int td,bool ft,bool m2){
  switch(tv){
    case B2S_VISCO_ARTIFICIAL: return p2_pick_dens<TKER,B2S_VISCO_ARTIFICIAL>(td,ft,m2);
    case B2S_VISCO_LAMINARSPS: return p2_pick_dens<TKER,B2S_VISCO_LAMINARSPS>(td,ft,m2);
    default:                   return p2_pick_dens<TKER,B2S_VISCO_LAMINAR>(td,ft,m2);
  }
}

int b2s_forces_pair2_launch(b2s_ctx* ctx, const ForcesArgs& g){
  const DevCte& c=ctx->cte;
  const uint32_t npf=ctx->npf_host;
  const unsigned gb=b2s_blocks(npf,P2_BS);
  const PackCte k=pack_cte(c);
  const bool ft=(c.use_floating!=0), m2=(c.mdbc2 && g.boundmode!=nullptr);
  const pair2_kernel_t kern=(c.tkernel==B2S_KERNEL_WENDLAND? p2_pick_visco<B2S_KERNEL_WENDLAND>(c.tvisco,c.tdensity,ft,m2): p2_pick_visco<B2S_KERNEL_CUBIC>(c.tvisco,c.tdensity,ft,m2));
  { static const char* cv=getenv("B2S_P2_CARVEOUT"); if(cv)cudaFuncSetAttribute((const void*)kern,cudaFuncAttributePreferredSharedMemoryCarveout,atoi(cv)); }   // experiment: share of the SM's memory taken from L1
  kern<<<gb,P2_BS,0,ctx->stream>>>(g,k,c);
  B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
