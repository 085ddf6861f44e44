// forces_tile.cu -- interaction variant 6: bricks of cells staged in shared memory, per-lane hit lists, packed-FP32 pair body.
//
// Same physics as variant 0 (forces.cu / forces_common.cuh pair_fluid), which restates cusph::Interaction_Forces /
// KerInteractionForcesFluid (src/source/JSphGpu_ker.cu:539-990) and JSphCpu::InteractionForcesFluid (src/source/JSphCpu.cpp:736-1125);
// the neighbour search replaces cunsearch::InitCte/ParticleRange (src/source/JCellSearch_iker.h:34-57,155-167).
//
// Why (profiles/README.md, round 1): the thread-per-particle kernels spend 24% of their instructions on per-lane row set-up (three
// dependent begin-table loads per row and lane), gather every candidate through L1 (8.1 KB of L1 traffic per particle) and lose a
// quarter of the pair-body rounds to the lane with most hits in a row.  Here
//   * a persistent block takes one BRICK of cells (3x2x2 kernel radii: 6x4x4 cells of size h or 3x2x2 cells of size 2h) from a
//     list of the non-empty bricks, and stages the brick plus one kernel radius of halo ONCE in shared memory: every neighbour row
//     (fixed y,z) is one contiguous index range of the cell-sorted arrays, so a tile is 64 (16) coalesced row copies per table;
//   * positions are converted on the way in to ONE frame per tile (FP64 difference to the brick centre, rounded once: the
//     precision of the reference's PosCell frame, JSphGpuSimple_ker.cu:46-57), so a pair needs no cell offsets: dr = pos_i - pos_j;
//   * the per-cell begin table of the tile lives in shared memory (two 16-bit loads per row and lane, after trimming the row to
//     the cells the kernel sphere reaches);
//   * every lane first checks its candidates (exact FP32 test, one funnel shift per candidate collects the hit bits), pushes
//     (mask, base) entries to a private list in shared memory, and only then evaluates the hits two at a time in f32x2 registers:
//     the lanes of a warp are decoupled from rows, a round is lost only to the lane with most hits in ~16 rows together.
// Per pair the operations are those of variants 4/5 (pair2_body); the sums are grouped differently and dr comes from the tile
// frame, so results agree with variant 0 to rounding (tests/test_gpu_parity.py).
// A tile that does not fit (more than T_CAP particles in the halo region: piled-up particles) is split along x, and a single
// column that still does not fit is evaluated from global memory with the variant-0 walk: no configuration is refused for density.
#include "forces_pair2.cuh"
#include <cuda_fp16.h>

#ifndef T_BS
#define T_BS 576             // threads per block = targets per pass over a brick (a full brick of 96 half cells holds ~500 particles)
#endif
#define T_NW (T_BS/32)
// Register budget: an SM sub-partition holds 16 K registers, so 18 warps (5 on two of the four sub-partitions) leave 96 per
// thread and 16 warps (T_BS 512) leave 128.  T_MAXREG selects __maxnreg__ instead of __launch_bounds__ (they exclude each other).
#ifdef T_MAXREG
#define T_KATTR __maxnreg__(T_MAXREG)
#else
#define T_KATTR __launch_bounds__(T_BS,1)
#endif
#define T_CAP 4096           // particles per tile (halo region 10x8x8 half cells = 3330 at the lattice density)
#define T_E 12               // list entries per lane

template<int D> struct TG {   // D = scelldiv (cells per kernel radius): same physical brick for both cell modes
  static constexpr int H=D, BX=3*D, BY=2*D, BZ=2*D, HX=BX+2*D, HY=BY+2*D, HZ=BZ+2*D, NROW=HY*HZ, NTROW=BY*BZ, W=HX+1, NR=(2*D+1)*(2*D+1);
};

struct TileArgs {
  const double2* posxy; const double* posz;
  const uint32_t* bricks;      // list of non-empty bricks (k_tile_bricks)
  uint32_t* counters;          // [0] bricks in the list, [1] next brick to take
  uint32_t cap;                // <= T_CAP (tests lower it to exercise the split and the global fallback)
  PackCte k;
};

// ---- list of non-empty bricks ------------------------------------------------------------------------------------------
template<int D> __global__ void __launch_bounds__(256) k_tile_bricks(const DivState* __restrict__ ds, const uint32_t* __restrict__ beg, uint32_t* __restrict__ list, uint32_t* counters){
  typedef TG<D> G;
  const int ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const int nbx=(ncx+G::BX-1)/G::BX, nby=(ncy+G::BY-1)/G::BY, nbz=(ncz+G::BZ-1)/G::BZ;
  const uint32_t nb=uint32_t(nbx)*uint32_t(nby)*uint32_t(nbz);
  const uint32_t* bt=beg+ds->box_fluid;
  const uint32_t nround=(nb+blockDim.x*gridDim.x-1)/(blockDim.x*gridDim.x);
  for(uint32_t it=0;it<nround;it++){
    const uint32_t b=(it*gridDim.x+blockIdx.x)*blockDim.x+threadIdx.x;
    uint32_t cnt=0;
    if(b<nb){
      const int bz=int(b/uint32_t(nbx*nby)), by=int((b-uint32_t(bz)*uint32_t(nbx*nby))/uint32_t(nbx)), bx=int(b)-bz*nbx*nby-by*nbx;
      const int x0=bx*G::BX, x1=min(x0+G::BX,ncx);
      for(int tz=0;tz<G::BZ;tz++)for(int ty=0;ty<G::BY;ty++){
        const int y=by*G::BY+ty, z=bz*G::BZ+tz;
        if(y<ncy && z<ncz){ const uint32_t* r=bt+(y*ncx+z*nsheet); cnt+=r[x1]-r[x0]; }
      }
    }
    // warp-aggregated append: the list stays (almost) in brick order, which keeps neighbouring bricks close in time (L2 reuse)
    const uint32_t bal=__ballot_sync(0xffffffffu,cnt!=0);
    if(bal){
      const int lane=threadIdx.x&31;
      uint32_t base=0;
      if(lane==0)base=atomicAdd(&counters[0],uint32_t(__popc(bal)));
      base=__shfl_sync(0xffffffffu,base,0);
      if(cnt)list[base+__popc(bal&((1u<<lane)-1u))]=b;
    }
  }
}

// ---- explicit shared-memory accesses (32-bit shared-window addresses: no generic-address arithmetic inside the loops) ----------
__device__ __forceinline__ float4 lds128(uint32_t a){ float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];":"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"r"(a)); return v; }
__device__ __forceinline__ uint4 lds128u(uint32_t a){ uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"r"(a)); return v; }
__device__ __forceinline__ uint2 lds64u(uint32_t a){ uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1},[%2];":"=r"(v.x),"=r"(v.y):"r"(a)); return v; }
// FP16 pre-filter of two candidates: t = dz*dz+(dy*dy+(dx*dx-thr)) in half2; sets bit0/bit1 of mask where t<0
__device__ __forceinline__ void hit2h(uint32_t& mask,uint32_t tx,uint32_t ty,uint32_t tz,uint32_t cx,uint32_t cy,uint32_t cz,uint32_t mthr,uint32_t bit0,uint32_t bit1){
  asm("{\n .reg .b32 dx,dy,dz,t; .reg .pred p,q;\n sub.rn.f16x2 dx,%1,%4;\n sub.rn.f16x2 dy,%2,%5;\n sub.rn.f16x2 dz,%3,%6;\n fma.rn.f16x2 t,dx,dx,%7;\n fma.rn.f16x2 t,dy,dy,t;\n fma.rn.f16x2 t,dz,dz,t;\n"
      " setp.lt.f16x2 p|q,t,%10;\n @p or.b32 %0,%0,%8;\n @q or.b32 %0,%0,%9;\n}"
      :"+r"(mask):"r"(tx),"r"(ty),"r"(tz),"r"(cx),"r"(cy),"r"(cz),"r"(mthr),"r"(bit0),"r"(bit1),"r"(0u));
}
__device__ __forceinline__ uint32_t lds32(uint32_t a){ uint32_t v; asm volatile("ld.shared.u32 %0,[%1];":"=r"(v):"r"(a)); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a){ uint16_t v; asm volatile("ld.shared.u16 %0,[%1];":"=h"(v):"r"(a)); return uint32_t(v); }
__device__ __forceinline__ void sts32(uint32_t a,uint32_t v){ asm volatile("st.shared.u32 [%0],%1;"::"r"(a),"r"(v):"memory"); }
__device__ __forceinline__ uint32_t opaque(uint32_t x){ asm volatile("mov.u32 %0,%1;":"=r"(x):"r"(x)); return x; }
__device__ __forceinline__ uint32_t bfind(uint32_t x){ uint32_t f; asm("bfind.u32 %0,%1;":"=r"(f):"r"(x)); return f; }
__device__ __forceinline__ void sts16(uint32_t a,uint32_t v){ asm volatile("st.shared.u16 [%0],%1;"::"r"(a),"h"(uint16_t(v)):"memory"); }

// ---- the kernel ----------------------------------------------------------------------------------------------------------
// Shared memory (dynamic), in this order:
//   tpos  float4[T_CAP]   position in the tile frame + pressure    tvel float4[T_CAP]   velocity + density
//   hxy   half[T_CAP/4][8]      hz half[T_CAP/4][4]                 FP16 copy of the positions in cell units, four particles per
//                                                                  group: x0..x3,y0..y3 / z0..z3 (pre-filter of the candidate walk)
//   emask u32[T_NW][T_E][32]   ebase u32[T_NW][T_E][32]            hit lists: 32 candidates per entry (mask, address)
//   tbeg  u16[2][NROW][W]      first tile index of every cell of the tile (fluid table, boundary table)
//   rowg  u32[2*NROW]          global index of the first particle of every row       rowoff u32[2*NROW+1]  tile index of the same
//   toff  u32[NTROW+1]         prefix of the target counts per target row
template<int D> struct TileSmem {
  typedef TG<D> G;
  static constexpr size_t o_tpos=0, o_tvel=o_tpos+size_t(T_CAP)*16, o_hxy=o_tvel+size_t(T_CAP)*16, o_hz=o_hxy+size_t(T_CAP)*4, o_emask=o_hz+size_t(T_CAP)*2, o_ebase=o_emask+size_t(T_NW)*T_E*32*4,
    o_tbeg=o_ebase+size_t(T_NW)*T_E*32*4, o_rowg=o_tbeg+((size_t(2)*G::NROW*G::W*2+15)/16)*16,
    o_rowoff=o_rowg+size_t(2)*G::NROW*4, o_toff=o_rowoff+size_t(2*G::NROW+4)*4, o_ctl=o_toff+size_t(G::NTROW+4)*4, total=o_ctl+64;
};

template<int TDENS,int D>
__global__ void T_KATTR k_forces_fluid_tile(ForcesArgs g, TileArgs ta, const __grid_constant__ DevCte c)
{
  typedef TG<D> G; typedef TileSmem<D> S;
  extern __shared__ __align__(16) unsigned char smem[];
  float4* tpos=reinterpret_cast<float4*>(smem+S::o_tpos);
  float4* tvel=reinterpret_cast<float4*>(smem+S::o_tvel);
  uint16_t* tbeg=reinterpret_cast<uint16_t*>(smem+S::o_tbeg);
  uint32_t* rowg=reinterpret_cast<uint32_t*>(smem+S::o_rowg);
  uint32_t* rowoff=reinterpret_cast<uint32_t*>(smem+S::o_rowoff);
  uint32_t* toff=reinterpret_cast<uint32_t*>(smem+S::o_toff);
  uint32_t* ctl=reinterpret_cast<uint32_t*>(smem+S::o_ctl);

  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5;
  const DivState* ds=g.ds;
  const int ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const int nbx=(ncx+G::BX-1)/G::BX, nby=(ncy+G::BY-1)/G::BY;
  const uint32_t nlist=ta.counters[0];
  const uint32_t* btab[2]={g.beg+ds->box_fluid,g.beg};   // table 0: fluid (+floating) neighbours, table 1: boundary neighbours
  // 32-bit shared-window addresses for the explicit ld.shared/st.shared of the pair loops
  const uint32_t s_base=uint32_t(__cvta_generic_to_shared(smem));
  const uint32_t s_tpos=opaque(s_base+uint32_t(S::o_tpos)), s_tvel=s_tpos+uint32_t(T_CAP)*16u, s_tbeg=s_base+uint32_t(S::o_tbeg);
  const uint32_t s_hxy=opaque(s_base+uint32_t(S::o_hxy)), s_hz=s_hxy+uint32_t(T_CAP)*4u;
  const float ovs=1.0f/c.scell;
  // this lane's list: entry e at +e*128 in both arrays (mask; shared address of candidate 31 of the chunk).  opaque(): the values
  // stay in registers (nvcc otherwise re-derives them from threadIdx.x inside the pair loop: 15 instructions per round)
  const uint32_t lm=opaque(s_base+uint32_t(S::o_emask)+uint32_t(warp*(T_E*32)+lane)*4u);
  const uint32_t lb=opaque(s_base+uint32_t(S::o_ebase)+uint32_t(warp*(T_E*32)+lane)*4u);
  float ace2=0.f, viscmax=0.f;

  for(;;){
    __syncthreads();                                       // the tile and ctl of the previous brick are no longer read
    if(tid==0)ctl[0]=atomicAdd(&ta.counters[1],1u);
    __syncthreads();
    const uint32_t bi=ctl[0];
    if(bi>=nlist)break;
    const uint32_t b=ta.bricks[bi];
    const int bz=int(b/uint32_t(nbx*nby)), by=int((b-uint32_t(bz)*uint32_t(nbx*nby))/uint32_t(nbx)), bx=int(b)-bz*nbx*nby-by*nbx;
    const int bx0=bx*G::BX, by0=by*G::BY, bz0=bz*G::BZ;
    // tile frame: origin at the centre of the brick (absolute domain cells), FP64 like the cell hash (JSphGpu_ker.cu:387-389)
    const double orgx=c.domposmin[0]+c.scell_d*double(int(ds->cellmin[0])+bx0+G::BX/2);
    const double orgy=c.domposmin[1]+c.scell_d*double(int(ds->cellmin[1])+by0+G::BY/2);
    const double orgz=c.domposmin[2]+c.scell_d*double(int(ds->cellmin[2])+bz0+G::BZ/2);

    int xs=0, w=G::BX;                                     // current column [bx0+xs, bx0+xs+w) of the brick
    while(xs<G::BX){
      if(bx0+xs>=ncx)break;
      const int xorg=bx0+xs-G::H;                          // first cell of the halo region along x (may be <0)
      const int hxn=w+2*G::H;                              // cells of the halo region along x
      __syncthreads();
      // 1. row ranges of both tables
      if(tid<2*G::NROW){
        const int tb=tid/G::NROW, r=tid-tb*G::NROW, hy=r%G::HY, hz=r/G::HY;
        const int y=by0-G::H+hy, z=bz0-G::H+hz;
        uint32_t gb=0, ge=0;
        if(y>=0 && y<ncy && z>=0 && z<ncz){
          const uint32_t* bt=btab[tb]+(y*ncx+z*nsheet);
          gb=bt[max(xorg,0)]; ge=bt[min(xorg+hxn,ncx)];
        }
        rowg[tid]=gb; rowoff[tid]=ge-gb;
      }
      __syncthreads();
      // 2. exclusive prefix of the row lengths (fluid rows first, then boundary rows)
      if(warp==0){
        constexpr int PER=(2*G::NROW+31)/32;
        uint32_t v[PER], sum=0;
        #pragma unroll
        for(int k=0;k<PER;k++){ const int i=lane*PER+k; v[k]=(i<2*G::NROW? rowoff[i]: 0u); sum+=v[k]; }
        uint32_t inc=sum;
        #pragma unroll
        for(int o=1;o<32;o<<=1){ const uint32_t t=__shfl_up_sync(0xffffffffu,inc,o); if(lane>=o)inc+=t; }
        uint32_t run=inc-sum;
        #pragma unroll
        for(int k=0;k<PER;k++){ const int i=lane*PER+k; if(i<2*G::NROW)rowoff[i]=run; run+=v[k]; }
        if(lane==31)rowoff[2*G::NROW]=inc;
      }
      __syncthreads();
      const uint32_t ntile=rowoff[2*G::NROW], nfl=rowoff[G::NROW];
      if(ntile>ta.cap){
        if(w>1){ w=(w+1)>>1; continue; }
        //-A single column that does not fit: its targets are evaluated from global memory (variant-0 walk, forces_common.cuh).
        for(int rt=0;rt<G::NTROW;rt++){
          const int ty=rt%G::BY, tz=rt/G::BY, y=by0+ty, z=bz0+tz;
          if(y>=ncy || z>=ncz)continue;
          const uint32_t* bt=btab[0]+(y*ncx+z*nsheet);
          const uint32_t pa=bt[bx0+xs], pb=bt[bx0+xs+1];
          for(uint32_t p1=pa+tid;p1<pb;p1+=T_BS){
            if(B2S_CODE_ISNORMAL(g.code[p1])){
              float a2=0.f,vm=0.f;
              gather_particle<B2S_KERNEL_WENDLAND,B2S_VISCO_ARTIFICIAL,TDENS,false>(g,c,ds,p1,true,a2,vm);
              ace2=fmaxf(ace2,a2); viscmax=fmaxf(viscmax,vm);
            }
          }
        }
        xs+=1; w=G::BX-xs;
        continue;
      }
      // 3. begin table of the tile: first tile index of every cell (columns 0..hxn), both tables
      for(int i=tid;i<2*G::NROW*G::W;i+=T_BS){
        const int rr=i/G::W, hx=i-rr*G::W, tb=rr/G::NROW, r=rr-tb*G::NROW, hy=r%G::HY, hz=r/G::HY;
        const int y=by0-G::H+hy, z=bz0-G::H+hz;
        uint32_t v=rowoff[rr];
        if(hx<=hxn && y>=0 && y<ncy && z>=0 && z<ncz){
          const uint32_t* bt=btab[tb]+(y*ncx+z*nsheet);
          const int x=min(max(xorg+hx,max(xorg,0)),min(xorg+hxn,ncx));
          v+=bt[x]-rowg[rr];
        }
        tbeg[i]=uint16_t(v);
      }
      // 4. particles: one row per warp and turn, FP64 position -> tile frame
      for(int rr=warp;rr<2*G::NROW;rr+=T_NW){
        const uint32_t gb=rowg[rr], o=rowoff[rr], n=rowoff[rr+1]-o;
        for(uint32_t k=lane;k<n;k+=32){
          const uint32_t p=gb+k;
          const double2 xy=ta.posxy[p]; const double z=ta.posz[p];
          float4 q; q.x=float(xy.x-orgx); q.y=float(xy.y-orgy); q.z=float(z-orgz); q.w=g.posp[p].w;
          tpos[o+k]=q; tvel[o+k]=g.velrho[p];
          const uint32_t sj=o+k, gh=s_hxy+(sj>>2)*16u+(sj&3u)*2u;
          sts16(gh,__half_as_ushort(__float2half_rn(q.x*ovs))); sts16(gh+8u,__half_as_ushort(__float2half_rn(q.y*ovs)));
          sts16(s_hz+(sj>>2)*8u+(sj&3u)*2u,__half_as_ushort(__float2half_rn(q.z*ovs)));
        }
      }
      __syncthreads();
      if(tid==0){
        uint32_t run=0;
        for(int rt=0;rt<G::NTROW;rt++){
          const int r=(rt%G::BY+G::H)+G::HY*(rt/G::BY+G::H);
          toff[rt]=run; run+=uint32_t(tbeg[r*G::W+G::H+w])-uint32_t(tbeg[r*G::W+G::H]);
        }
        toff[G::NTROW]=run;
      }
      __syncthreads();
      const uint32_t ntarget=toff[G::NTROW];
      const bool hasbound=(ntile>nfl);

      // 5. targets: T_BS at a time
      for(uint32_t t0=0;t0<ntarget;t0+=T_BS){
        if(t0+warp*32>=ntarget)continue;                   // whole warp without targets (warp-uniform)
        const uint32_t t=t0+tid;
        bool valid=(t<ntarget);
        uint32_t si=0, p1=0; int r0=G::H+G::HY*G::H, hxc=G::H;
        if(valid){
          int rt=0;
          #pragma unroll
          for(int k=1;k<G::NTROW;k++)if(t>=toff[k])rt=k;
          r0=(rt%G::BY+G::H)+G::HY*(rt/G::BY+G::H);
          si=uint32_t(tbeg[r0*G::W+G::H])+(t-toff[rt]);
          p1=rowg[r0]+(si-rowoff[r0]);
          #pragma unroll
          for(int k=1;k<G::BX;k++)if(k<w && si>=uint32_t(tbeg[r0*G::W+G::H+k]))hxc=G::H+k;   // the target's own cell (halo coordinates)
          valid=B2S_CODE_ISNORMAL(g.code[p1]);             // slab halo ghosts / periodic duplicates are neighbours only
        }
        const float4 pp1=lds128(s_tpos+si*16u), velrhop1=(valid? lds128(s_tvel+si*16u): make_float4(0,0,0,1.f));
        const float pressp1=pp1.w, ovrhop1=b2s_rcp(velrhop1.w);
        uint32_t tx2,ty2,tz2;                              // the target in the FP16 copy's units, both halves
        { const __half2 a=__float2half2_rn(pp1.x*ovs), b=__float2half2_rn(pp1.y*ovs), d=__float2half2_rn(pp1.z*ovs); tx2=*reinterpret_cast<const uint32_t*>(&a); ty2=*reinterpret_cast<const uint32_t*>(&b); tz2=*reinterpret_cast<const uint32_t*>(&d); }
        uint32_t mthr2; { const __half2 t=__float2half2_rn(-float(G::H*G::H)*1.0125f); mthr2=*reinterpret_cast<const uint32_t*>(&t); }
        const uint32_t selfadr=s_tpos+si*16u;         // list base of the target itself as the zero-contribution partner (see pack_body)
        PackAcc acc;
        pack_zero(acc);
        PackTarget tgt; tgt.tensil=0.f; tgt.ft=false;      // Wendland + artificial viscosity without floating bodies: unused
        PackNeigh nbx; nbx.mu=nbx.plain=pk1(1.f); nbx.heavy_a=nbx.heavy_b=false; nbx.taua=nbx.taub=nullptr;
        PairAcc sbnd={0,0,0,0,0,0,0,0,0,0,0,0};
        #pragma unroll 1
        for(int pass=(hasbound? 1: 0);pass>=0;pass--){
          // first row of the window of this lane in the table of this pass: cells [hxc-H, hxc+H] of rows r0-H-HY*H ...
          const uint32_t rbase=s_tbeg+uint32_t((pass*G::NROW+r0-G::H-G::HY*G::H)*G::W+hxc-G::H)*2u;
          uint32_t ne=0, nh=0;                             // entries / hits waiting in this lane's list
          #pragma unroll 1
          for(int rowi=0;rowi<=G::NR;rowi++){
            uint32_t sa=0, a0=0, pre=0, len=0, maxlen=0;
            if(rowi<G::NR){
              // Every lane walks the WHOLE window of its cell in this row (cells hxc-H..hxc+H): the lanes of one cell read the
              // same addresses (one shared-memory broadcast instead of 32 private streams); the exact test below drops the rest.
              const int jz=rowi/(2*G::H+1), jy=rowi-jz*(2*G::H+1);
              const uint32_t ra=rbase+uint32_t((jy+G::HY*jz)*G::W)*2u;
              if(valid){ sa=lds16(ra); a0=sa&~3u; pre=sa-a0; len=lds16(ra+uint32_t(2*G::H+1)*2u)-a0; if(len==pre)len=0u; }   // len counts from a0
              maxlen=__reduce_max_sync(0xffffffffu,len);
            }
            uint32_t c0=0;
            do{
              if(c0<maxlen){
                //-FP16 pre-filter of up to 32 candidates, four per step (aligned groups of the FP16 copy): rr2 in cell units
                // against H^2 with a margin above the FP16 error (1.25%: 2% more hits, which w=max(w,0) zeroes in pack_body)
                const uint32_t nmax=min(32u,maxlen-c0);
                const uint32_t g0=(a0+c0)>>2;
                const uint32_t qxy=s_hxy+g0*16u, qz=s_hz+g0*8u;
                uint32_t m=0, mb=0;                         // two masks: two independent chains of predicated ORs
                #pragma unroll
                for(uint32_t k0=0;k0<32u;k0+=4u){
                  if(k0<nmax){
                    const uint4 xy=lds128u(qxy+k0*4u); const uint2 zz=lds64u(qz+k0*2u);
                    hit2h(m,tx2,ty2,tz2,xy.x,xy.z,zz.x,mthr2,1u<<k0,2u<<k0);
                    hit2h(mb,tx2,ty2,tz2,xy.y,xy.w,zz.y,mthr2,4u<<k0,8u<<k0);
                  }
                }
                m|=mb;
                //-candidates of the chunk outside [sa,sb): the ones in front of sa in its group, the ones behind sb
                if(c0==0u)m&=(0xffffffffu<<pre);
                { const uint32_t nl=(len>c0? len-c0: 0u); if(nl<32u)m&=((1u<<nl)-1u); }
                if(m){ sts32(lm+ne*128u,m); sts32(lb+ne*128u,s_tpos+(a0+c0)*16u); ne++; nh+=uint32_t(__popc(m)); }   // base = address of candidate 0
              }
              c0+=32u;
              if(rowi==G::NR || __any_sync(0xffffffffu,ne>=uint32_t(T_E-3))){
                //-drain: every lane evaluates its own hits two at a time; a round is lost only to the lane with most hits.
                // An odd count is padded with the target itself (contributes 0).  (cm,cb) = entry being consumed.
                while(nh&3u){ sts32(lm+ne*128u,1u); sts32(lb+ne*128u,selfadr); ne++; nh++; }   // pad to a multiple of four with the target itself
                uint32_t pm=lm, pb_=lb, cm=0u, cb=0u;
                const uint32_t rounds=__reduce_max_sync(0xffffffffu,nh>>2);
                //-two hits leave the list: addresses of their tile entries
                auto pop2=[&](uint32_t& ja,uint32_t& jb){
                  if(cm==0u){ cm=lds32(pm); cb=lds32(pb_); pm+=128u; pb_+=128u; }
                  uint32_t f=bfind(cm); cm^=(1u<<f); ja=cb+(f<<4);
                  if(cm==0u){ cm=lds32(pm); cb=lds32(pb_); pm+=128u; pb_+=128u; }
                  f=bfind(cm); cm^=(1u<<f); jb=cb+(f<<4);
                };
                auto eval=[&](const float4& pa,const float4& pb,const float4& va,const float4& vb){
                  const f2 drx=pk(FSUB(pp1.x,pa.x),FSUB(pp1.x,pb.x)), dry=pk(FSUB(pp1.y,pa.y),FSUB(pp1.y,pb.y)), drz=pk(FSUB(pp1.z,pa.z),FSUB(pp1.z,pb.z));
                  const f2 rr2=fma2(drz,drz,fma2(dry,dry,mul2(drx,drx)));   // b2s_rr2 order
                  if(pass==0)pack_body<B2S_KERNEL_WENDLAND,B2S_VISCO_ARTIFICIAL,TDENS,false,false,false>(c,ta.k,drx,dry,drz,rr2,velrhop1,pressp1,ovrhop1,tgt,va,vb,pa.w,pb.w,nbx,acc);
                  else       pack_body<B2S_KERNEL_WENDLAND,B2S_VISCO_ARTIFICIAL,TDENS,false,false,true >(c,ta.k,drx,dry,drz,rr2,velrhop1,pressp1,ovrhop1,tgt,va,vb,pa.w,pb.w,nbx,acc);
                };
                //-four hits per turn: two independent packed evaluations in flight (the kernel runs at 4-5 warps per scheduler,
                // so the instruction-level parallelism has to come from inside the warp)
                #pragma unroll 1
                for(uint32_t it=0;it<rounds;it++){
                  if(4u*it<nh){
                    uint32_t ja,jb,jc,jd; pop2(ja,jb); pop2(jc,jd);
                    const float4 pa=lds128(ja), pb=lds128(jb), pc=lds128(jc), pd=lds128(jd);
                    const float4 va=lds128(ja+uint32_t(T_CAP)*16u), vb=lds128(jb+uint32_t(T_CAP)*16u), vc=lds128(jc+uint32_t(T_CAP)*16u), vd=lds128(jd+uint32_t(T_CAP)*16u);
                    eval(pa,pb,va,vb);
                    eval(pc,pd,vc,vd);
                  }
                }
                ne=0; nh=0;
              }
            }while(c0<maxlen);
          }
          if(pass==1){ // scale and keep the boundary sums, restart the accumulators for the fluid pass
            pack_finish<B2S_KERNEL_WENDLAND,B2S_VISCO_ARTIFICIAL,TDENS,true>(c,ta.k,acc,velrhop1.w,false,sbnd);
            pack_zero(acc);
          }
        }
        if(valid){
          PairAcc sf={0,0,0,0,0,0,0,0,0,0,0,0};
          pack_finish<B2S_KERNEL_WENDLAND,B2S_VISCO_ARTIFICIAL,TDENS,false>(c,ta.k,acc,velrhop1.w,false,sf);
          float a2=0.f,vm=0.f;
          finalize_store<B2S_VISCO_ARTIFICIAL,TDENS>(g,c,p1,velrhop1,sf,sbnd,a2,vm);
          ace2=fmaxf(ace2,a2); viscmax=fmaxf(viscmax,vm);
        }
      }
      xs+=w; w=G::BX-xs;
    }
  }
  //-Fused maxima: ViscDtMax and AceMax^2 (replaces ReduMaxFloat x2 + cudaMemcpy, JSphGpu_ker.cu:75-125): once per block.
  reduce_maxima(g.st,viscmax,ace2);
}

// ------------------------------------------------------------------------------------------------
bool b2s_forces_tile_supported(const DevCte& c){
  return c.tkernel==B2S_KERNEL_WENDLAND && c.tvisco==B2S_VISCO_ARTIFICIAL && !c.use_floating && !c.mdbc2 && (c.scelldiv==1 || c.scelldiv==2);
}

template<int D> static int tile_launch(b2s_ctx* ctx, const ForcesArgs& g){
  typedef TG<D> G;
  const DevCte& c=ctx->cte;
  if(!ctx->tile_bricks){
    const uint64_t nb=(uint64_t(c.domcells[0])/G::BX+2)*(uint64_t(c.domcells[1])/G::BY+2)*(uint64_t(c.domcells[2])/G::BZ+2);
    B2S_CUDA(ctx,cudaMalloc((void**)&ctx->tile_bricks,size_t(nb)*4+64));
    B2S_CUDA(ctx,cudaMalloc((void**)&ctx->tile_counters,16));
    int nsm=0; B2S_CUDA(ctx,cudaDeviceGetAttribute(&nsm,cudaDevAttrMultiProcessorCount,ctx->device)); ctx->tile_nsm=nsm;
    #define T_ATTR(TD) B2S_CUDA(ctx,cudaFuncSetAttribute(k_forces_fluid_tile<TD,D>,cudaFuncAttributeMaxDynamicSharedMemorySize,int(TileSmem<D>::total)));
    T_ATTR(B2S_DDT_NONE) T_ATTR(B2S_DDT_MOLTENI) T_ATTR(B2S_DDT_FOURTAKAS) T_ATTR(B2S_DDT_FOURTAKASFULL)
    #undef T_ATTR
  }
  cudaStream_t s=ctx->stream;
  B2S_CUDA(ctx,cudaMemsetAsync(ctx->tile_counters,0,16,s));
  const uint64_t nbmax=(uint64_t(ctx->nct_max)/(G::BX*G::BY*G::BZ))+1;
  k_tile_bricks<D><<<unsigned(std::min<uint64_t>((nbmax+255)/256,uint64_t(ctx->tile_nsm)*8)),256,0,s>>>(ctx->div,(const uint32_t*)ctx->begincell,ctx->tile_bricks,ctx->tile_counters); B2S_LAUNCHED(ctx);
  TileArgs ta; ta.posxy=ctx->cur.posxy; ta.posz=ctx->cur.posz; ta.bricks=ctx->tile_bricks; ta.counters=ctx->tile_counters;
  ta.cap=(ctx->tile_cap && ctx->tile_cap<T_CAP? ctx->tile_cap: T_CAP);
  ta.k=pack_cte(c);
  const unsigned gb=unsigned(ctx->tile_nsm);
  const size_t sm=TileSmem<D>::total;
  switch(c.tdensity){
    case B2S_DDT_NONE:      k_forces_fluid_tile<B2S_DDT_NONE,D><<<gb,T_BS,sm,s>>>(g,ta,c); break;
    case B2S_DDT_MOLTENI:   k_forces_fluid_tile<B2S_DDT_MOLTENI,D><<<gb,T_BS,sm,s>>>(g,ta,c); break;
    case B2S_DDT_FOURTAKAS: k_forces_fluid_tile<B2S_DDT_FOURTAKAS,D><<<gb,T_BS,sm,s>>>(g,ta,c); break;
    default:                k_forces_fluid_tile<B2S_DDT_FOURTAKASFULL,D><<<gb,T_BS,sm,s>>>(g,ta,c); break;
  }
  B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}

int b2s_forces_tile_launch(b2s_ctx* ctx, const ForcesArgs& g){
  return ctx->cte.scelldiv==2? tile_launch<2>(ctx,g): tile_launch<1>(ctx,g);
}
