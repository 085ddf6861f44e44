// forces_tile.cu -- variant 1 of the fluid interaction kernel: packed-FP16 coarse filter + exact FP32 pair body.
//
// Why (measured on B200, tools/microbench.cu, profiles/): the reference-style gather loop (variant 0,
// KerInteractionForcesFluid, src/source/JSphGpu_ker.cu:786-990) spends one 128-bit load per *candidate* per warp and
// only ~15% of the candidates are inside the kernel radius, while the divergent hit test makes every lane pay for the
// full pair body.  A 128-bit shared/L1 load costs 4 LSU cycles per warp regardless of broadcast (128 B/clk/SM return
// path), so even a shared-memory staged FP32 scan is LSU-bound.  This kernel therefore splits the work:
//   1. stage   : one CTA owns T_NCX x-adjacent target cells; the particles of the (T_NCX+2) x 3 x 3 neighbour cells
//                (fluid rows, then boundary rows) are staged ONCE as FP16 coordinates relative to the segment
//                (6 B per candidate), 8 candidates per 16-byte group.
//   2. scan    : each lane owns one target particle; per group the warp issues 3 broadcast LDS.128 and 24 packed
//                half2 ops (HADD2/HMUL2/HFMA2, two candidates per instruction) and keeps an 8-bit hit mask per lane.
//                The FP16 test uses a widened radius (margin derived below) so it never misses a real neighbour.
//   3. body    : every lane walks ITS OWN hit bits (flattened over the mask words, so lanes stay busy although their
//                hits differ), gathers posp/velrho of the hit from global memory (L1-resident neighbourhood) and runs
//                the exact FP32 test + pair body of forces_common.cuh.  Hits are visited in increasing candidate
//                order = the reference's summation order, so variant 1 is bit-identical to variant 0.
// FP16 margin: coordinates are in units of the cell size (= 2h), centred on the segment, |u|<2.5 -> spacing 2^-9,
// rounding error <= 2^-10 per coordinate per particle; |d(dx)| <= 2^-9 + 2^-11; at r~1: |d(r2)| <= 2*sqrt(3)*0.0025 + 3*2^-11*1.1
// < 0.011.  The filter accepts r2_fp16 <= 1.03 (>2.5x the bound); the exact test then decides.
#include "forces_common.cuh"
#include <cuda_fp16.h>

#define T_NCX    2                 // target cells per segment
#define T_WARPS  4
#define T_BS     (T_WARPS*32)
#define T_KMAX   2560              // staged candidates per segment (multiple of 8)
#define T_GMAX   (T_KMAX/8)
#define T_WGMAX  240               // groups scanned per warp round (1920 candidates), multiple of 4
#define T_NROWS  18                // 9 fluid rows + 9 boundary rows
#define T_THR2   1.03f

struct TileSmem {
  uint4 hx[T_GMAX], hy[T_GMAX], hz[T_GMAX];    // 8 halves per group
  uint32_t gp2[T_GMAX];                        // particle index of the first candidate of the group
  uint8_t cofs[T_KMAX];                        // (cx-x0+1) | row<<2
  uint32_t wmask[T_WARPS][T_WGMAX/4][32];      // per warp: mask word (4 groups) per lane
  uint16_t glist[T_WARPS][T_WGMAX];
  uint32_t rowk[T_NROWS+1];                    // first staged candidate of each row (padded to 8)
  uint32_t rowp[T_NROWS];                      // first particle of each row
  uint32_t rown[T_NROWS];                      // particles in each row
  int      rowxa[T_NROWS];                     // first cell x of each row range
  uint32_t tbeg, tend, kfluid, ktotal;
};

// ---- compact list of non-empty segments (built after every cell divide) ------------------------------------------
__global__ void __launch_bounds__(256) k_tile_seglist(const DivState* __restrict__ ds, const uint32_t* __restrict__ begt, uint32_t* seglist, uint32_t* nseg){
  const uint32_t ncx=ds->ncx, ncy=ds->ncy, ncz=ds->ncz;
  const uint32_t nxs=(ncx+T_NCX-1)/T_NCX;
  const uint32_t total=nxs*ncy*ncz;
  const uint32_t* beg=begt+ds->box_fluid;
  for(uint32_t base=blockIdx.x*blockDim.x; base<total; base+=gridDim.x*blockDim.x){
    const uint32_t s=base+threadIdx.x;
    bool ne=false;
    if(s<total){
      const uint32_t xs=s%nxs, row=s/nxs;
      const uint32_t x0=xs*T_NCX, x1=min(x0+T_NCX,ncx);
      const uint32_t c0=x0+row*ncx;
      ne=(beg[c0+(x1-x0)]>beg[c0]);
    }
    const uint32_t bal=__ballot_sync(0xffffffffu,ne);
    uint32_t wbase=0;
    if((threadIdx.x&31)==0 && bal)wbase=atomicAdd(nseg,__popc(bal));
    wbase=__shfl_sync(0xffffffffu,wbase,0);
    if(ne)seglist[wbase+__popc(bal&((1u<<(threadIdx.x&31))-1u))]=s;
  }
}

__device__ __forceinline__ uint32_t h2_mask_le(__half2 a,__half2 b){ return __hle2_mask(a,b); }

// scan of one round: glist[0..n) groups (n multiple of 4) -> wmask words
__device__ __forceinline__ void tile_scan(const TileSmem& sm,const uint16_t* __restrict__ glist,uint32_t n,uint32_t (*wmask)[32],int lane,
  __half2 tx,__half2 ty,__half2 tz,__half2 thr)
{
  for(uint32_t i0=0;i0<n;i0+=4){
    uint32_t word=0;
    #pragma unroll
    for(int q=0;q<4;q++){
      const uint32_t g=glist[i0+q];
      const uint4 X=sm.hx[g], Y=sm.hy[g], Z=sm.hz[g];
      uint32_t t=0;
      #define TILE_PAIR(XX,YY,ZZ,BITS) { \
        const __half2 dx=__hsub2(tx,*reinterpret_cast<const __half2*>(&XX)); \
        const __half2 dy=__hsub2(ty,*reinterpret_cast<const __half2*>(&YY)); \
        const __half2 dz=__hsub2(tz,*reinterpret_cast<const __half2*>(&ZZ)); \
        const __half2 r2=__hfma2(dz,dz,__hfma2(dy,dy,__hmul2(dx,dx))); \
        t|=(h2_mask_le(r2,thr)&(BITS)); }
      TILE_PAIR(X.x,Y.x,Z.x,0x00020001u)
      TILE_PAIR(X.y,Y.y,Z.y,0x00080004u)
      TILE_PAIR(X.z,Y.z,Z.z,0x00200010u)
      TILE_PAIR(X.w,Y.w,Z.w,0x00800040u)
      #undef TILE_PAIR
      const uint32_t byte=(t|(t>>16))&0xffu;
      word|=byte<<(8*q);
    }
    wmask[i0>>2][lane]=word;
  }
}

template<int TKER,int TVISCO,int TDENS,bool FLOATING,bool BOUNDP2>
__device__ __forceinline__ void tile_body(const TileSmem& sm,const ForcesArgs& g,const DevCte& c,const uint16_t* __restrict__ glist,uint32_t nwords,
  const uint32_t (*wmask)[32],int lane,const float4 pp1,const float4 velrhop1,bool ftp1,const float* taup1,int tcx,PairAcc& a)
{
  const float visco=(BOUNDP2? c.visco_b: c.visco_f);
  uint32_t wi=0;
  uint32_t m=(nwords? wmask[0][lane]: 0u);
  for(;;){
    while(m==0u && wi+1<nwords){ wi++; m=wmask[wi][lane]; }
    const bool has=(m!=0u);
    if(!__any_sync(0xffffffffu,has))break;
    if(has){
      const int b=__ffs(m)-1;
      m&=(m-1u);
      const uint32_t grp=glist[wi*4+(b>>3)];
      const uint32_t bit=uint32_t(b&7);
      const uint32_t p2=sm.gp2[grp]+bit;
      const uint32_t co=sm.cofs[grp*8+bit];
      const int dcx=int(co&3u)-1;                 // candidate cell x relative to the segment's first cell
      const int ri=int(co>>2)%9;
      const int dy=ri%3-1, dz=ri/3-1;             // candidate row relative to the target row
      const float4 pp2=g.posp[p2];
      const float drx=b2s_dr(pp1.x,pp2.x,c.scell*float(tcx-dcx));
      const float dry=b2s_dr(pp1.y,pp2.y,c.scell*float(-dy));
      const float drz=b2s_dr(pp1.z,pp2.z,c.scell*float(-dz));
      const float rr2=b2s_rr2(drx,dry,drz);
      if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO){
        const typecode cod2=((FLOATING && !BOUNDP2)? g.code[p2]: typecode(0));
        pair_fluid<TKER,TVISCO,TDENS,FLOATING,BOUNDP2>(c,drx,dry,drz,rr2,velrhop1,pp1.w,ftp1,taup1,
          g.velrho[p2],pp2.w,cod2,g.tau,p2,g.ftomassp,visco,a);
      }
    }
  }
}

template<int TKER,int TVISCO,int TDENS,bool FLOATING>
__global__ void __launch_bounds__(T_BS,4) k_forces_fluid_tile(ForcesArgs g,const uint32_t* __restrict__ seglist,const uint32_t* __restrict__ nsegp,
  const __grid_constant__ DevCte c)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  TileSmem& sm=*reinterpret_cast<TileSmem*>(smraw);
  const DivState* ds=g.ds;
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5;
  const uint32_t nseg=*nsegp;
  const int ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const int nxs=(ncx+T_NCX-1)/T_NCX;
  const float ovscell=1.0f/c.scell;
  float ace2=0.f, viscmax=0.f;
  for(uint32_t si=blockIdx.x; si<nseg; si+=gridDim.x){
    const uint32_t s=seglist[si];
    const int xs=int(s%uint32_t(nxs)), rowi=int(s/uint32_t(nxs));
    const int y=rowi%ncy, z=rowi/ncy;
    const int x0=xs*T_NCX, x1=min(x0+T_NCX,ncx);
    __syncthreads();   // previous segment fully consumed
    // ---- row table (one warp)
    if(warp==0){
      uint32_t n=0,p=0; int xa=0;
      if(lane<T_NROWS){
        const int pass=lane/9, ri=lane%9, dy=ri%3-1, dz=ri/3-1;
        const int yy=y+dy, zz=z+dz;
        if(yy>=0 && yy<ncy && zz>=0 && zz<ncz){
          xa=max(x0-1,0); const int xb=min(x1+1,ncx);
          const uint32_t* beg=g.beg+(pass==0? ds->box_fluid: 0u);
          const int rb=yy*ncx+zz*nsheet;
          p=beg[rb+xa]; n=beg[rb+xb]-p;
        }
        sm.rowp[lane]=p; sm.rown[lane]=n; sm.rowxa[lane]=xa;
      }
      uint32_t pad=(lane<T_NROWS? ((n+7u)&~7u): 0u);
      uint32_t incl=pad;
      #pragma unroll
      for(int o=1;o<32;o<<=1){ const uint32_t v=__shfl_up_sync(0xffffffffu,incl,o); if(lane>=o)incl+=v; }
      if(lane<T_NROWS)sm.rowk[lane]=incl-pad;
      if(lane==T_NROWS-1)sm.rowk[T_NROWS]=incl;
      if(lane==8)sm.kfluid=incl;
      if(lane==T_NROWS-1)sm.ktotal=incl;
      if(lane==0){
        const uint32_t* begf=g.beg+ds->box_fluid;
        const int c0=x0+y*ncx+z*nsheet;
        sm.tbeg=begf[c0]; sm.tend=begf[c0+(x1-x0)];
      }
    }
    __syncthreads();
    const uint32_t ktotal=sm.ktotal;
    if(ktotal>T_KMAX-8){
      // neighbourhood does not fit the staging buffer (pathological compression): flag and skip; the host reports it.
      if(tid==0)atomicCAS(&g.st->error,0,B2S_ERR_UNSUPPORTED);
      continue;
    }
    // ---- stage FP16 coordinates
    {
      __half* hx=reinterpret_cast<__half*>(sm.hx); __half* hy=reinterpret_cast<__half*>(sm.hy); __half* hz=reinterpret_cast<__half*>(sm.hz);
      const uint32_t kfill=ktotal+8;   // + one all-pad group used to pad the group lists
      for(uint32_t k=tid;k<kfill;k+=T_BS){
        int r=0;
        #pragma unroll
        for(int q=1;q<T_NROWS;q++)if(k>=sm.rowk[q])r=q;
        const uint32_t j=k-sm.rowk[r];
        float ux=100.f,uy=100.f,uz=100.f; uint32_t co=0;
        if(k<ktotal && j<sm.rown[r]){
          const uint32_t p2=sm.rowp[r]+j;
          const float4 pp=g.posp[p2];
          const int cx=int(B2S_DCEL_CELLX(c.domcellcode,g.dcell[p2]))-int(ds->cellmin[0]);
          const int ri=r%9, dy=ri%3-1, dz=ri/3-1;
          ux=float(cx-x0)+pp.x*ovscell-0.5f*float(T_NCX);
          uy=float(dy)+pp.y*ovscell-0.5f;
          uz=float(dz)+pp.z*ovscell-0.5f;
          co=uint32_t(cx-x0+1)|(uint32_t(r)<<2);
        }
        hx[k]=__float2half_rn(ux); hy[k]=__float2half_rn(uy); hz[k]=__float2half_rn(uz);
        sm.cofs[k]=uint8_t(co);
        if((k&7u)==0u)sm.gp2[k>>3]=sm.rowp[r]+j;
      }
    }
    __syncthreads();
    const uint32_t tbeg=sm.tbeg, tend=sm.tend;
    const uint32_t padgroup=ktotal>>3;   // the all-pad group
    // ---- warp tasks: 32 consecutive targets each
    for(uint32_t tw=warp; tw*32u<tend-tbeg; tw+=T_WARPS){
      const uint32_t p1=tbeg+tw*32u+lane;
      const bool active=(p1<tend);
      float4 pp1=make_float4(0,0,0,0), velrhop1=make_float4(0,0,0,1.f);
      int tcx=0; bool ftp1=false;
      float taup1[6]={0,0,0,0,0,0};
      float ux=-100.f,uy=-100.f,uz=-100.f;
      if(active){
        pp1=g.posp[p1]; velrhop1=g.velrho[p1];
        tcx=int(B2S_DCEL_CELLX(c.domcellcode,g.dcell[p1]))-int(ds->cellmin[0])-x0;
        if(FLOATING)ftp1=B2S_CODE_ISFLOATING(g.code[p1]);
        if(TVISCO==B2S_VISCO_LAMINARSPS){
          const float2* t=(const float2*)(g.tau+size_t(p1)*6);
          const float2 t0=t[0],t1=t[1],t2=t[2];
          taup1[0]=t0.x; taup1[1]=t0.y; taup1[2]=t1.x; taup1[3]=t1.y; taup1[4]=t2.x; taup1[5]=t2.y;
        }
        ux=float(tcx)+pp1.x*ovscell-0.5f*float(T_NCX); uy=pp1.y*ovscell-0.5f; uz=pp1.z*ovscell-0.5f;
      }
      const __half2 tx=__float2half2_rn(ux), ty=__float2half2_rn(uy), tz=__float2half2_rn(uz);
      const __half2 thr=__float2half2_rn(T_THR2);
      const int cxmin=__reduce_min_sync(0xffffffffu,active? tcx: 1000);
      const int cxmax=__reduce_max_sync(0xffffffffu,active? tcx: -1000);
      PairAcc af={0,0,0,0,0,0,0,0,0,0,0,0};
      PairAcc ab={0,0,0,0,0,0,0,0,0,0,0,0};
      if(FLOATING && ftp1 && TDENS!=B2S_DDT_NONE){ af.delta=FLT_MAX; ab.delta=FLT_MAX; }
      uint16_t* glist=sm.glist[warp];
      uint32_t (*wmask)[32]=sm.wmask[warp];
      // rounds over the rows: fluid rows (pass 0) then boundary rows (pass 1); a round ends when the group list is full
      for(int pass=0;pass<2;pass++){
        int r=pass*9; const int rend=r+9;
        uint32_t gcur=0xffffffffu, gend=0;   // pending group range of the current row
        for(;;){
          uint32_t n=0;
          while(n<T_WGMAX){
            if(gcur>=gend){
              if(r>=rend)break;
              // cells needed from row r: x in [x0+cxmin-1, x0+cxmax+1] clipped to the staged range
              const uint32_t nrow=sm.rown[r];
              gcur=gend=0;
              if(nrow){
                const int ri=r%9, dy=ri%3-1, dz=ri/3-1;
                const int rb=(y+dy)*ncx+(z+dz)*nsheet;
                const uint32_t* beg=g.beg+(pass==0? ds->box_fluid: 0u);
                const int xa=sm.rowxa[r];
                const int xl=max(x0+cxmin-1,xa), xh=min(x0+cxmax+1,min(x1+1,ncx)-1);
                const uint32_t ka=sm.rowk[r]+(beg[rb+xl]-sm.rowp[r]);
                const uint32_t kb=sm.rowk[r]+(beg[rb+xh+1]-sm.rowp[r]);
                gcur=ka>>3; gend=(kb+7u)>>3;
              }
              r++;
              continue;
            }
            const uint32_t take=min(gend-gcur,uint32_t(T_WGMAX)-n);
            for(uint32_t q=lane;q<take;q+=32)glist[n+q]=uint16_t(gcur+q);
            n+=take; gcur+=take;
          }
          if(n==0)break;
          const uint32_t npad=(n+3u)&~3u;
          if(lane<int(npad-n))glist[n+lane]=uint16_t(padgroup);
          __syncwarp();
          tile_scan(sm,glist,npad,wmask,lane,tx,ty,tz,thr);
          __syncwarp();
          if(pass==0)tile_body<TKER,TVISCO,TDENS,FLOATING,false>(sm,g,c,glist,npad>>2,wmask,lane,pp1,velrhop1,ftp1,taup1,tcx,af);
          else       tile_body<TKER,TVISCO,TDENS,FLOATING,true >(sm,g,c,glist,npad>>2,wmask,lane,pp1,velrhop1,ftp1,taup1,tcx,ab);
          __syncwarp();
          if(gcur>=gend && r>=rend)break;
        }
      }
      if(active){
        float a2=0.f,vm=0.f;
        finalize_store<TVISCO,TDENS>(g,c,p1,velrhop1,af,ab,a2,vm);
        ace2=fmaxf(ace2,a2); viscmax=fmaxf(viscmax,vm);
      }
    }
  }
  reduce_maxima(g.st,viscmax,ace2);
}

// ---- dispatch ---------------------------------------------------------------------------------------------------------
typedef void (*tile_kernel_t)(ForcesArgs,const uint32_t*,const uint32_t*,const DevCte);
template<int TKER,int TVISCO,int TDENS> static tile_kernel_t tpick_ft(bool ft){
  return ft? (tile_kernel_t)k_forces_fluid_tile<TKER,TVISCO,TDENS,true>: (tile_kernel_t)k_forces_fluid_tile<TKER,TVISCO,TDENS,false>;
}
template<int TKER,int TVISCO> static tile_kernel_t tpick_dens(int td,bool ft){
  switch(td){
    case B2S_DDT_NONE: return tpick_ft<TKER,TVISCO,B2S_DDT_NONE>(ft);
    case B2S_DDT_MOLTENI: return tpick_ft<TKER,TVISCO,B2S_DDT_MOLTENI>(ft);
    case B2S_DDT_FOURTAKAS: return tpick_ft<TKER,TVISCO,B2S_DDT_FOURTAKAS>(ft);
    default: return tpick_ft<TKER,TVISCO,B2S_DDT_FOURTAKASFULL>(ft);
  }
}
template<int TKER> static tile_kernel_t tpick_visco(int tv,int td,bool ft){
  switch(tv){
    case B2S_VISCO_ARTIFICIAL: return tpick_dens<TKER,B2S_VISCO_ARTIFICIAL>(td,ft);
    case B2S_VISCO_LAMINARSPS: return tpick_dens<TKER,B2S_VISCO_LAMINARSPS>(td,ft);
    default: return tpick_dens<TKER,B2S_VISCO_LAMINAR>(td,ft);
  }
}

int b2s_tile_build_seglist(b2s_ctx* ctx){
  if(!ctx->tile_seglist){
    const uint64_t nmax=uint64_t((ctx->prm.domcells[0]+T_NCX-1)/T_NCX)*ctx->prm.domcells[1]*ctx->prm.domcells[2];
    B2S_CUDA(ctx,cudaMalloc((void**)&ctx->tile_seglist,size_t(nmax+1)*4));
    B2S_CUDA(ctx,cudaMalloc((void**)&ctx->tile_nseg,64));
    ctx->tile_nseg_max=uint32_t(nmax);
  }
  B2S_CUDA(ctx,cudaMemsetAsync(ctx->tile_nseg,0,4,ctx->stream));
  const unsigned gb=std::min<unsigned>(b2s_blocks(ctx->tile_nseg_max,256),148u*16u);
  k_tile_seglist<<<gb,256,0,ctx->stream>>>(ctx->div,(const uint32_t*)ctx->begincell,ctx->tile_seglist,ctx->tile_nseg); B2S_LAUNCHED(ctx);
  return B2S_OK;
}

int b2s_forces_tile_launch(b2s_ctx* ctx, const ForcesArgs& g){
  const DevCte& c=ctx->cte;
  if(c.scelldiv!=1)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"tile interaction variant needs cellmode full (cells of 2h)");
  const bool ft=(c.use_floating!=0);
  tile_kernel_t k=(c.tkernel==B2S_KERNEL_WENDLAND? tpick_visco<B2S_KERNEL_WENDLAND>(c.tvisco,c.tdensity,ft): tpick_visco<B2S_KERNEL_CUBIC>(c.tvisco,c.tdensity,ft));
  const size_t smem=sizeof(TileSmem);
  if(!ctx->tile_attr_set[ft?1:0][c.tkernel&1][c.tvisco&3][c.tdensity&3]){
    B2S_CUDA(ctx,cudaFuncSetAttribute((const void*)k,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
    ctx->tile_attr_set[ft?1:0][c.tkernel&1][c.tvisco&3][c.tdensity&3]=true;
  }
  if(!ctx->tile_seglist_valid){ int r=b2s_tile_build_seglist(ctx); if(r)return r; ctx->tile_seglist_valid=true; }
  int occ=0;
  B2S_CUDA(ctx,cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,(const void*)k,T_BS,smem));
  if(occ<1)occ=1;
  const unsigned grid=148u*unsigned(occ);
  k<<<grid,T_BS,smem,ctx->stream>>>(g,ctx->tile_seglist,ctx->tile_nseg,ctx->cte); B2S_LAUNCHED(ctx);
  return B2S_OK;
}
