// forces_defer.cu -- interaction variant 3: gather with DEFERRED pair bodies (sm_100a).
//
// Same physics, same neighbour order and therefore bit-identical results to variant 0 (forces.cu), which restates
// cusph::Interaction_Forces / KerInteractionForcesFluid (src/source/JSphGpu_ker.cu:673-964) and the CPU loop
// JSphCpu::InteractionForcesFluid (src/source/JSphCpu.cpp:736-1125).
//
// Why: the kernel is FP32-issue bound (profiles/r01_*: 80% issue-active, 3.2 IPC) and in variant 0 the ~70-instruction
// pair body runs under the `rr2<4h^2` branch with only ~27% of the lanes active (half cells; 16% with 2h cells): a warp
// pays the body for nearly every candidate because at least one of its 32 lanes almost always hits.
// Here one thread still owns one target particle and walks the same candidates in the same order, but a hit is only
// PUSHED (index + 6-bit row code) into a per-lane FIFO in shared memory; bodies are executed in rounds in which every
// lane pops one entry, so the expensive code runs with (nearly) all 32 lanes active.  Rounds are taken when every lane
// has an entry (full efficiency); a lane whose FIFO could overflow forces a round with the lanes that have entries.
//   * FIFO: Q_CAP x 32 lanes x 4 B per warp, slot-major so that push and pop are conflict-free;
//   * a row of neighbour cells is ONE index range thanks to the block-relative x of posp (b2s_internal.h, B2S_XBLK):
//     the x offset is one select (p2<pmid) instead of a per-cell inner loop;
//   * the row code travels with the entry because the body recomputes dr exactly like the check did.
#include "forces_common.cuh"

#define D_BS    128          // threads per block (4 warps)
#ifndef Q_CAP
#define Q_CAP   16           // FIFO slots per lane (power of two).  Small on purpose: shared memory is carved out of L1,
#endif                       // and the scattered loads of the deferred bodies live on L1 hits (profiles/r01_defer_tuning.md)
#ifndef Q_CHUNK
#define Q_CHUNK 4            // candidates checked between two drain decisions
#endif
#ifndef D_MINB
#define D_MINB 6
#endif
#define D_SMEM  ((D_BS/32)*Q_CAP*32*4 + 32*8)
#define Q_BYTES (Q_CAP*128)  // one lane's FIFO column spans Q_CAP slots of 128 B (32 lanes x 4 B)

__device__ __forceinline__ void sts32(uint32_t addr,uint32_t v){ asm volatile("st.shared.u32 [%0], %1;"::"r"(addr),"r"(v):"memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr){ uint32_t v; asm volatile("ld.shared.u32 %0, [%1];":"=r"(v):"r"(addr):"memory"); return v; }
__device__ __forceinline__ float2 lds64f(uint32_t addr){ float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];":"=f"(v.x),"=f"(v.y):"r"(addr)); return v; }
// predicated 16-byte load through the read-only path (posp is never written while the interaction kernel runs)
__device__ __forceinline__ float4 ldg128_if(const float4* p,bool pred){
  float4 v;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %5, 0;\n @q ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n}"
               :"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"l"(p),"r"(int(pred)));
  return v;
}

template<int TKER,int TVISCO,int TDENS,bool FLOATING>
__global__ void __launch_bounds__(D_BS,D_MINB) k_forces_fluid_defer(ForcesArgs g, const __grid_constant__ DevCte c)
{
  extern __shared__ __align__(16) uint32_t dsm[];
  const uint32_t sbase=uint32_t(__cvta_generic_to_shared(dsm));        // [32] float2 (offy,offz) per row code, then the FIFOs
  // this lane's FIFO column (32-bit shared address): slot s lives at qbase + s*128
  const uint32_t qbase=sbase+256u+(threadIdx.x>>5)*Q_BYTES+(threadIdx.x&31)*4u;
  const DivState* ds=g.ds;
  const uint32_t npb=ds->npb, np=ds->np;
  const uint32_t p1=npb+blockIdx.x*blockDim.x+threadIdx.x;
  const bool valid=(p1<np);
  const int d=c.scelldiv, nrow=2*d+1;
  if(threadIdx.x<32){
    const int rz=int(threadIdx.x)/nrow, ry=int(threadIdx.x)-rz*nrow;    // row code = (z-(cz-d))*nrow + (y-(cy-d))
    reinterpret_cast<float2*>(dsm)[threadIdx.x]=make_float2(c.scell*float(d-ry),c.scell*float(d-rz));
  }
  __syncthreads();
  float ace2=0.f, viscmax=0.f;
  // ---- target particle
  float4 pp1=make_float4(0,0,0,0), velrhop1=make_float4(0,0,0,1.f);
  bool ftp1=false;
  float taup1[6]={0,0,0,0,0,0};
  CellRange r; r.cx=r.cy=r.cz=0; r.xini=r.xfin=r.yini=r.yfin=r.zini=r.zfin=0;
  if(valid){
    pp1=g.posp[p1]; velrhop1=g.velrho[p1];
    if(FLOATING)ftp1=B2S_CODE_ISFLOATING(g.code[p1]);
    if(TVISCO==B2S_VISCO_LAMINARSPS){
      const float2* t=(const float2*)(g.tau+size_t(p1)*6);
      const float2 t0=t[0],t1=t[1],t2=t[2];
      taup1[0]=t0.x; taup1[1]=t0.y; taup1[2]=t1.x; taup1[3]=t1.y; taup1[4]=t2.x; taup1[5]=t2.y;
    }
    r=cell_range(g.dcell[p1],c,ds);
  }
  const float pressp1=pp1.w;
  const int ncx=int(ds->ncx), nsheet=int(ds->nsheet), cmx=int(ds->cellmin[0]);
  // x of the target inside its own cell (posp.x is block-relative) -- used to trim the rows, never in the pair arithmetic
  const float xin=pp1.x-c.scell*float((r.cx+cmx)&((1<<B2S_XBLK)-1));
  // the untrimmed x window touches at most two blocks of 2^B2S_XBLK cells: A (of xini) and B (of xfin-1)
  const bool oneblk=(((r.xini+cmx)>>B2S_XBLK)==((r.xfin-1+cmx)>>B2S_XBLK));
  const int xmid0=(((r.xfin-1+cmx)>>B2S_XBLK)<<B2S_XBLK)-cmx;           // first cell of block B
  const float sxA=FADD(pp1.x,b2s_offx(c,r.cx+cmx,r.xini+cmx)), sxB=FADD(pp1.x,b2s_offx(c,r.cx+cmx,r.xfin-1+cmx));
  PairAcc af={0,0,0,0,0,0,0,0,0,0,0,0};
  PairAcc ab={0,0,0,0,0,0,0,0,0,0,0,0};
  if(FLOATING && ftp1 && TDENS!=B2S_DDT_NONE){ af.delta=FLT_MAX; ab.delta=FLT_MAX; }

  uint32_t hoff=0, toff=0, cnt=0;      // FIFO of this lane: byte offsets of head/tail slot (mod Q_BYTES) and entry count
  #pragma unroll 1
  for(int pass=0;pass<2;pass++){
    const uint32_t* beg=(pass==0? g.beg+ds->box_fluid: g.beg);   // pass 0: fluid/floating neighbours, pass 1: boundary
    // n rounds: in each one every lane that has an entry pops it and runs the pair body
    auto rounds=[&](uint32_t n){
      #pragma unroll 1
      for(;n;n--)if(cnt){
        const uint32_t e=lds32(qbase+hoff); hoff=(hoff+128u)&(Q_BYTES-1u); cnt--;
        const uint32_t p2=e&0x3ffffffu;
        const float2 off=lds64f(sbase+((e>>23)&0xf8u));
        const float4 pp2=__ldg(g.posp+p2);
        const float drx=FSUB(int(e)<0? sxB: sxA,pp2.x), dry=FSUB(FADD(pp1.y,off.x),pp2.y), drz=FSUB(FADD(pp1.z,off.y),pp2.z);
        const float rr2=b2s_rr2(drx,dry,drz);
        if(pass==0){
          const typecode cod2=(FLOATING? g.code[p2]: typecode(0));
          pair_fluid<TKER,TVISCO,TDENS,FLOATING,false>(c,drx,dry,drz,rr2,velrhop1,pressp1,ftp1,taup1,__ldg(g.velrho+p2),pp2.w,cod2,g.tau,p2,g.ftomassp,c.visco_f,af);
        }
        else pair_fluid<TKER,TVISCO,TDENS,FLOATING,true>(c,drx,dry,drz,rr2,velrhop1,pressp1,ftp1,taup1,__ldg(g.velrho+p2),pp2.w,typecode(0),g.tau,p2,g.ftomassp,c.visco_b,ab,g.boundmode,g.tangenvel);
      }
    };
    #pragma unroll 1
    for(int z=r.cz-d;z<=r.cz+d;z++){
      const float offz=c.scell*float(r.cz-z);
      const float sz=FADD(pp1.z,offz);
      #pragma unroll 1
      for(int y=r.cy-d;y<=r.cy+d;y++){
        const float offy=c.scell*float(r.cy-y);
        const float sy=FADD(pp1.y,offy);
        uint32_t pbeg=0,pend=0,pmid=0;
        int xa,xb;
        if(z>=r.zini && z<r.zfin && y>=r.yini && y<r.yfin && row_trim(c,r,xin,sy,sz,xa,xb)){
          const uint32_t* brow=beg+(y*ncx+z*nsheet);
          pbeg=brow[xa]; pend=brow[xb];
          pmid=(oneblk || xmid0<=xa? pbeg: (xmid0>=xb? pend: brow[xmid0]));
        }
        const uint32_t codeA=uint32_t((z-r.cz+d)*nrow+(y-r.cy+d))<<26, codeB=codeA|0x80000000u;
        const uint32_t lenmax=__reduce_max_sync(0xffffffffu,pend-pbeg);
        #pragma unroll 1
        for(uint32_t k0=0;k0<lenmax;k0+=Q_CHUNK){
          //-Before a chunk: as many rounds as every lane can take part in (full efficiency), and at least as many as the
          // fullest FIFO needs to have room for a whole chunk (forced: only the lanes with entries work).
          {
            const uint32_t mx=__reduce_max_sync(0xffffffffu,cnt), mn=__reduce_min_sync(0xffffffffu,valid? cnt: uint32_t(Q_CAP));
            rounds(max(mn,max(mx,uint32_t(Q_CAP-Q_CHUNK))-uint32_t(Q_CAP-Q_CHUNK)));
          }
          //-Check one chunk of candidates, push the hits.
          #pragma unroll
          for(uint32_t k=0;k<Q_CHUNK;k++){
            const uint32_t p2=pbeg+k0+k;
            const bool in=(p2<pend);
            const float4 pp2=ldg128_if(g.posp+p2,in);
            const bool inb=(p2>=pmid);
            const float drx=FSUB(inb? sxB: sxA,pp2.x), dry=FSUB(sy,pp2.y), drz=FSUB(sz,pp2.z);
            const float rr2=b2s_rr2(drx,dry,drz);
            if(in && rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO){ sts32(qbase+toff,p2|(inb? codeB: codeA)); toff=(toff+128u)&(Q_BYTES-1u); cnt++; }
          }
        }
      }
    }
    //-End of the pass: flush what is left (the two passes use different bodies).
    rounds(__reduce_max_sync(0xffffffffu,cnt));
  }
  if(valid)finalize_store<TVISCO,TDENS>(g,c,p1,velrhop1,af,ab,ace2,viscmax);
  //-Fused maxima: ViscDtMax and AceMax^2 (replaces ReduMaxFloat x2 + cudaMemcpy, JSphGpu_ker.cu:75-125).
  #pragma unroll
  for(int o=16;o;o>>=1){ viscmax=fmaxf(viscmax,__shfl_xor_sync(0xffffffffu,viscmax,o)); ace2=fmaxf(ace2,__shfl_xor_sync(0xffffffffu,ace2,o)); }
  if((threadIdx.x&31)==0){
    if(viscmax>0.f)atomicMax(&g.st->viscdtmax_bits,__float_as_uint(viscmax));
    if(ace2>0.f)atomicMax(&g.st->acemax2_bits,__float_as_uint(ace2));
  }
}

// ------------------------------------------------------------------------------------------------
typedef void (*defer_kernel_t)(ForcesArgs, const DevCte);
template<int TKER,int TVISCO,int TDENS> static defer_kernel_t dpick_ft(bool ft){
  return ft? (defer_kernel_t)k_forces_fluid_defer<TKER,TVISCO,TDENS,true>: (defer_kernel_t)k_forces_fluid_defer<TKER,TVISCO,TDENS,false>;
}
template<int TKER,int TVISCO> static defer_kernel_t dpick_dens(int td,bool ft){
  switch(td){
    case B2S_DDT_NONE: return dpick_ft<TKER,TVISCO,B2S_DDT_NONE>(ft);
    case B2S_DDT_MOLTENI: return dpick_ft<TKER,TVISCO,B2S_DDT_MOLTENI>(ft);
    case B2S_DDT_FOURTAKAS: return dpick_ft<TKER,TVISCO,B2S_DDT_FOURTAKAS>(ft);
    default: return dpick_ft<TKER,TVISCO,B2S_DDT_FOURTAKASFULL>(ft);
  }
}
template<int TKER> static defer_kernel_t dpick_visco(int tv,int td,bool ft){
  switch(tv){
    case B2S_VISCO_ARTIFICIAL: return dpick_dens<TKER,B2S_VISCO_ARTIFICIAL>(td,ft);
    case B2S_VISCO_LAMINARSPS: return dpick_dens<TKER,B2S_VISCO_LAMINARSPS>(td,ft);
    default: return dpick_dens<TKER,B2S_VISCO_LAMINAR>(td,ft);
  }
}

int b2s_forces_defer_launch(b2s_ctx* ctx, const ForcesArgs& g){
  const DevCte& c=ctx->cte;
  const int nrow=2*c.scelldiv+1;
  if(nrow*nrow>32)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"interaction variant 3 supports cellmode full/half only");
  const bool ft=(c.use_floating!=0);
  defer_kernel_t k=(c.tkernel==B2S_KERNEL_WENDLAND? dpick_visco<B2S_KERNEL_WENDLAND>(c.tvisco,c.tdensity,ft): dpick_visco<B2S_KERNEL_CUBIC>(c.tvisco,c.tdensity,ft));
  const uint32_t npf=ctx->np_host-ctx->npb_host;
  k<<<b2s_blocks(npf,D_BS),D_BS,D_SMEM,ctx->stream>>>(g,c); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
