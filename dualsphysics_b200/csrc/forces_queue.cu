// forces_queue.cu -- variant 2 of the fluid interaction kernel: warp-uniform candidate stream + per-lane hit queue.
//
// Measured on B200 (profiles/r01_ncu_forces_variants.md): the reference-style gather loop (variant 0) is ISSUE bound
// (80% issue-slot utilisation) at 16 of 32 lanes active, because the ~100-instruction pair body runs under a divergent
// `if(rr2<=4h^2)` that only ~15% of the candidates pass, and because lanes of one warp walk different cell ranges.
// This kernel keeps the reference's memory behaviour (no staging; positions/velocities are read through a full-size
// L1, candidate loads are warp-uniform = broadcast) but restructures the control flow:
//   * the warp walks ONE candidate stream: the 3x3 rows around the warp's row, x-cells [cxmin-1, cxmax+1] (lanes whose
//     cell is further away simply fail the distance test) -> loop control and candidate loads are uniform;
//   * the per-candidate work is only the exact FP32 distance test (+ a predicated 4-byte push into the lane's own
//     queue in shared memory);
//   * when a queue is nearly full the warp "flushes": every lane pops ITS OWN hits and runs the pair body with all
//     lanes busy.  Hits are processed in stream order = the reference's summation order, so results are bit-identical
//     to variant 0.
// Lanes whose particles lie in another (y,z) row or more than one cell further in x are handled in a later group
// iteration (rare: row ends, sparse free-surface cells).
#include "forces_common.cuh"

#define Q_BS    128
#define Q_CAP   64          // queue entries per lane
#define Q_CHUNK 16          // candidates scanned between two overflow checks
#define Q_P2MASK 0x03ffffffu   // 26 bits particle index + 6 bits cell-offset code
#define Q_P2BITS 26

struct QCtx {
  float4 pp1, velrhop1; bool ftp1; float taup1[6];
};

template<int TKER,int TVISCO,int TDENS,bool FLOATING,bool BOUNDP2>
__device__ __forceinline__ void queue_flush(uint32_t (*q)[Q_BS],int tid,uint32_t& cnt,const ForcesArgs& g,const DevCte& c,
  const QCtx& t,int cx,int lcx,PairAcc& a)
{
  const uint32_t maxc=__reduce_max_sync(0xffffffffu,cnt);
  const float visco=(BOUNDP2? c.visco_b: c.visco_f);
  for(uint32_t k=0;k<maxc;k++){
    if(k<cnt){
      const uint32_t e=q[k][tid];
      const uint32_t p2=e&Q_P2MASK;
      const int code=int(e>>Q_P2BITS);           // (dxc+1) | (dy+1)<<2 | (dz+1)<<4, dxc = candidate cell x - lcx in [-1,2]
      const int dxc=(code&3)-1, dy=((code>>2)&3)-1, dz=(code>>4)-1;
      const float4 pp2=g.posp[p2];
      const float drx=b2s_dr(t.pp1.x,pp2.x,c.scell*float(cx-(lcx+dxc)));
      const float dry=b2s_dr(t.pp1.y,pp2.y,c.scell*float(-dy));
      const float drz=b2s_dr(t.pp1.z,pp2.z,c.scell*float(-dz));
      const float rr2=b2s_rr2(drx,dry,drz);
      if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO){   // exact test, same expression as variant 0 (the scan test is 1e-5 wider)
        const typecode cod2=((FLOATING && !BOUNDP2)? g.code[p2]: typecode(0));
        pair_fluid<TKER,TVISCO,TDENS,FLOATING,BOUNDP2>(c,drx,dry,drz,rr2,t.velrhop1,t.pp1.w,t.ftp1,t.taup1,
          g.velrho[p2],pp2.w,cod2,g.tau,p2,g.ftomassp,visco,a);
      }
    }
  }
  cnt=0;
}

template<int TKER,int TVISCO,int TDENS,bool FLOATING>
__global__ void __launch_bounds__(Q_BS) k_forces_fluid_queue(ForcesArgs g, const __grid_constant__ DevCte c)
{
  __shared__ uint32_t q[Q_CAP][Q_BS];
  const DivState* ds=g.ds;
  const int tid=threadIdx.x;
  const uint32_t npb=ds->npb, np=ds->np;
  const uint32_t p1=npb+blockIdx.x*blockDim.x+tid;
  const bool valid=(p1<np);
  const int ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  QCtx t; t.pp1=make_float4(0,0,0,0); t.velrhop1=make_float4(0,0,0,1.f); t.ftp1=false;
  #pragma unroll
  for(int k=0;k<6;k++)t.taup1[k]=0.f;
  int cx=0,cy=0,cz=0;
  if(valid){
    t.pp1=g.posp[p1]; t.velrhop1=g.velrho[p1];
    const uint32_t dc=g.dcell[p1];
    cx=int(B2S_DCEL_CELLX(c.domcellcode,dc))-int(ds->cellmin[0]);
    cy=int(B2S_DCEL_CELLY(c.domcellcode,dc))-int(ds->cellmin[1]);
    cz=int(B2S_DCEL_CELLZ(c.domcellcode,dc))-int(ds->cellmin[2]);
    if(FLOATING)t.ftp1=B2S_CODE_ISFLOATING(g.code[p1]);
    if(TVISCO==B2S_VISCO_LAMINARSPS){
      const float2* tt=(const float2*)(g.tau+size_t(p1)*6);
      const float2 t0=tt[0],t1=tt[1],t2=tt[2];
      t.taup1[0]=t0.x; t.taup1[1]=t0.y; t.taup1[2]=t1.x; t.taup1[3]=t1.y; t.taup1[4]=t2.x; t.taup1[5]=t2.y;
    }
  }
  PairAcc af={0,0,0,0,0,0,0,0,0,0,0,0};
  PairAcc ab={0,0,0,0,0,0,0,0,0,0,0,0};
  if(FLOATING && t.ftp1 && TDENS!=B2S_DDT_NONE){ af.delta=FLT_MAX; ab.delta=FLT_MAX; }
  const float ks2w=c.kernelsize2*1.00001f;
  bool todo=valid;
  for(;;){
    const uint32_t tb=__ballot_sync(0xffffffffu,todo);
    if(!tb)break;
    const int leader=__ffs(tb)-1;
    const int lcx=__shfl_sync(0xffffffffu,cx,leader), lcy=__shfl_sync(0xffffffffu,cy,leader), lcz=__shfl_sync(0xffffffffu,cz,leader);
    const bool ing=(todo && cy==lcy && cz==lcz && cx>=lcx && cx<=lcx+1);
    const int gmax=__reduce_max_sync(0xffffffffu,ing? cx: lcx);
    const int xlo=max(lcx-1,0), xhi=min(gmax+1,ncx-1);
    // per-lane position relative to the leader cell (lanes outside the group are moved far away: they never hit)
    const float bx=(ing? t.pp1.x+c.scell*float(cx-lcx): 1e18f), by=t.pp1.y, bz=t.pp1.z;
    #pragma unroll
    for(int pass=0;pass<2;pass++){
      const uint32_t* beg=g.beg+(pass==0? ds->box_fluid: 0u);
      uint32_t cnt=0;
      for(int dz=-1;dz<=1;dz++){
        const int zz=lcz+dz; if(zz<0 || zz>=ncz)continue;
        for(int dy=-1;dy<=1;dy++){
          const int yy=lcy+dy; if(yy<0 || yy>=ncy)continue;
          const int rb=yy*ncx+zz*nsheet;
          const float ty=by-c.scell*float(dy), tz=bz-c.scell*float(dz);
          for(int x=xlo;x<=xhi;x++){
            const uint32_t pini=beg[rb+x], pfin=beg[rb+x+1];
            if(pini==pfin)continue;
            // a lane only interacts with its own 3x3x3 cells (like the reference): cells two columns away are masked out,
            // otherwise a lattice pair at exactly r=2h across a whole cell could add a FLT_MAX boundary marker the reference never sees
            const float tx=((x-cx<=1 && cx-x<=1)? bx-c.scell*float(x-lcx): 1e18f);
            const uint32_t code=uint32_t((x-lcx+1)|((dy+1)<<2)|((dz+1)<<4))<<Q_P2BITS;
            for(uint32_t pc=pini;pc<pfin;pc+=Q_CHUNK){
              if(__any_sync(0xffffffffu,cnt>Q_CAP-Q_CHUNK)){
                if(pass==0)queue_flush<TKER,TVISCO,TDENS,FLOATING,false>(q,tid,cnt,g,c,t,cx,lcx,af);
                else       queue_flush<TKER,TVISCO,TDENS,FLOATING,true >(q,tid,cnt,g,c,t,cx,lcx,ab);
              }
              const uint32_t pe=min(pc+Q_CHUNK,pfin);
              for(uint32_t pb=pc;pb<pe;pb+=4){
                // 4 independent uniform loads in flight, then 4 tests (tail candidates are clamped and masked)
                float4 pp[4];
                #pragma unroll
                for(int u=0;u<4;u++)pp[u]=g.posp[min(pb+u,pe-1u)];
                #pragma unroll
                for(int u=0;u<4;u++){
                  const float dx=tx-pp[u].x, dy_=ty-pp[u].y, dz_=tz-pp[u].z;
                  const float rr2=dx*dx+dy_*dy_+dz_*dz_;
                  if(rr2<=ks2w && pb+u<pe){ q[cnt][tid]=(pb+u)|code; cnt++; }
                }
              }
            }
          }
        }
      }
      if(pass==0)queue_flush<TKER,TVISCO,TDENS,FLOATING,false>(q,tid,cnt,g,c,t,cx,lcx,af);
      else       queue_flush<TKER,TVISCO,TDENS,FLOATING,true >(q,tid,cnt,g,c,t,cx,lcx,ab);
    }
    todo=(todo && !ing);
  }
  float ace2=0.f, viscmax=0.f;
  if(valid)finalize_store<TVISCO,TDENS>(g,c,p1,t.velrhop1,af,ab,ace2,viscmax);
  reduce_maxima(g.st,viscmax,ace2);
}

typedef void (*queue_kernel_t)(ForcesArgs,const DevCte);
template<int TKER,int TVISCO,int TDENS> static queue_kernel_t qpick_ft(bool ft){
  return ft? (queue_kernel_t)k_forces_fluid_queue<TKER,TVISCO,TDENS,true>: (queue_kernel_t)k_forces_fluid_queue<TKER,TVISCO,TDENS,false>;
}
template<int TKER,int TVISCO> static queue_kernel_t qpick_dens(int td,bool ft){
  switch(td){
    case B2S_DDT_NONE: return qpick_ft<TKER,TVISCO,B2S_DDT_NONE>(ft);
    case B2S_DDT_MOLTENI: return qpick_ft<TKER,TVISCO,B2S_DDT_MOLTENI>(ft);
    case B2S_DDT_FOURTAKAS: return qpick_ft<TKER,TVISCO,B2S_DDT_FOURTAKAS>(ft);
    default: return qpick_ft<TKER,TVISCO,B2S_DDT_FOURTAKASFULL>(ft);
  }
}
template<int TKER> static queue_kernel_t qpick_visco(int tv,int td,bool ft){
  switch(tv){
    case B2S_VISCO_ARTIFICIAL: return qpick_dens<TKER,B2S_VISCO_ARTIFICIAL>(td,ft);
    case B2S_VISCO_LAMINARSPS: return qpick_dens<TKER,B2S_VISCO_LAMINARSPS>(td,ft);
    default: return qpick_dens<TKER,B2S_VISCO_LAMINAR>(td,ft);
  }
}

int b2s_forces_queue_launch(b2s_ctx* ctx, const ForcesArgs& g){
  const DevCte& c=ctx->cte;
  if(c.scelldiv!=1)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"queue interaction variant needs cellmode full (cells of 2h)");
  if(ctx->cap>Q_P2MASK)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"queue interaction variant supports up to 2^26 particles per GPU");
  const bool ft=(c.use_floating!=0);
  queue_kernel_t k=(c.tkernel==B2S_KERNEL_WENDLAND? qpick_visco<B2S_KERNEL_WENDLAND>(c.tvisco,c.tdensity,ft): qpick_visco<B2S_KERNEL_CUBIC>(c.tvisco,c.tdensity,ft));
  const uint32_t npf=(ctx->np_host>ctx->npb_host? ctx->np_host-ctx->npb_host: 0u);
  k<<<b2s_blocks(npf,Q_BS),Q_BS,0,ctx->stream>>>(g,ctx->cte); B2S_LAUNCHED(ctx);
  return B2S_OK;
}
