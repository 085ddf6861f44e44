// mdbc.cu -- modified Dynamic Boundary Condition: density (and, for mDBC2, velocity) extrapolated from the fluid to the ghost
// node of every boundary particle that carries a normal.
//
// Replaces cusph::Interaction_MdbcCorrection / Interaction_Mdbc2Correction (src/source/JSphGpu_mdbc_iker.cu:33-213,289-582),
// called from JSphGpuSingle::MdbcBoundCorrection (src/source/JSphGpuSingle.cpp:686-712); the reference's CPU statement of the
// same mathematics is JSphCpu::InteractionMdbcCorrectionT2 / InteractionMdbc2CorrectionT2 (src/source/JSphCpu_mdbc.cpp:36-150,
// 246-436).  What has to be computed per ghost node g (boundary particle + 2*normal, JSph.cpp:1417):
//     b  = sum_j m_j (W, gradW)                      Shepard density and its gradient (4 numbers)
//     A  = sum_j V_j (W, gradW) (x) (1, r_g - r_j)    the first-order consistency matrix (4x4), V_j = m_j/rho_j
//   then  x = A^-1 b:  rho_g = x0,  grad(rho)_g = -x[1..3]  (first order, if A is well conditioned), else Shepard b0/A00, else rho0.
//
// B200 design (the reference runs one thread per boundary particle over ~250 neighbours with 20 running sums):
//   * ONE WARP per ghost node.  Every row of neighbour cells (fixed y,z) is one contiguous range of the cell-sorted arrays, so the
//     32 lanes stride over the row: coalesced 128-bit loads of posp, the begin-table reads are warp-uniform broadcasts.
//   * The whole neighbour window is ONE dense candidate list: lane r reads the begin table of row r (one load latency for all 25
//     rows), a warp scan of the row lengths orders the candidates, and the lanes walk that list 32 at a time.
//   * Test first, evaluate later: the hits of the rows are compacted through a per-warp queue in shared memory, so the kernel
//     function and the FP64 accumulations run on 32 hits at a time (a quarter of the candidates are hits).
//   * Every lane keeps the 20 (mDBC2: 25) partial sums in FP64 registers; five xor-shuffle rounds leave the totals in all lanes.
//   * The 4x4 system is solved by Gaussian elimination with row pivoting on [A | b] (mDBC2: [A | I], its condition number needs
//     the inverse), FP64, fully unrolled in registers, uniform across the warp; the determinant the reference thresholds on is
//     the signed product of the pivots.
// The boundary normals arrive already doubled (boundary -> ghost node).  The pressure cached for the interaction (posp.w) is
// refreshed for the corrected density.
#include "b2s_internal.h"
#include <float.h>

#define M_BS 128                 // 4 ghost nodes per block
#define M_WARPS (M_BS/32)

// ---- neighbour window of a ghost node and the walk over it ---------------------------------------------------------------
struct GhostNode {
  float gx,gy,gz;                // position relative to the origin of its own cell (y,z) / of its block of cells (x): the posp frame
  int acx;                       // absolute cell x (selects the x block)
  int cx,cy,cz;                  // cell relative to the divide's cell domain
  int xini,xfin,yini,yfin,zini,zfin;
};

__device__ __forceinline__ GhostNode ghost_node(const DevCte& c,const DivState* ds,double px,double py,double pz,const float4 nor){
  GhostNode g;
  double x=px+double(nor.x), y=py+double(nor.y), z=pz+double(nor.z);
  if(c.periactive){ uint32_t cell; b2s_pos_to_cell(c,x,y,z,cell); }   // a ghost node behind a periodic face is wrapped (JSphCpu_mdbc.cpp:50,267)
  const double rx=x-c.domposmin[0], ry=y-c.domposmin[1], rz=z-c.domposmin[2];
  const int ax=int(rx/c.scell_d), ay=int(ry/c.scell_d), az=int(rz/c.scell_d);
  g.acx=ax;
  g.gx=float(rx-c.scell_d*double((ax>>B2S_XBLK)<<B2S_XBLK)); g.gy=float(ry-c.scell_d*double(ay)); g.gz=float(rz-c.scell_d*double(az));
  g.cx=ax-int(ds->cellmin[0]); g.cy=ay-int(ds->cellmin[1]); g.cz=az-int(ds->cellmin[2]);
  const int d=c.scelldiv, ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz);
  g.xini=g.cx-min(g.cx,d); g.xfin=g.cx+min(ncx-g.cx-1,d)+1;
  g.yini=g.cy-min(g.cy,d); g.yfin=g.cy+min(ncy-g.cy-1,d)+1;
  g.zini=g.cz-min(g.cz,d); g.zfin=g.cz+min(ncz-g.cz-1,d)+1;
  if(!(g.xfin>g.xini && g.xini>=0 && g.yini>=0 && g.zini>=0)){ g.zini=0; g.zfin=0; }   // node outside the cell domain: no neighbours
  return g;
}

// Partial sums of one lane.  s[0..3] = b, a[r][k] = A, and for mDBC2 the Shepard velocity sum, sum V W, and the divergence-of-position
// test value of the reference ("submerged").
template<bool MDBC2> struct GhostSums {
  double s[4]; double a[4][4]; double v[3]; double sub;
  __device__ __forceinline__ void zero(){
    #pragma unroll
    for(int r=0;r<4;r++){ s[r]=0;
      #pragma unroll
      for(int k=0;k<4;k++)a[r][k]=0; }
    v[0]=v[1]=v[2]=0; sub=0;
  }
  __device__ __forceinline__ void warp_total(){
    #pragma unroll
    for(int o=16;o;o>>=1){
      #pragma unroll
      for(int r=0;r<4;r++){ s[r]+=__shfl_xor_sync(0xffffffffu,s[r],o);
        #pragma unroll
        for(int k=0;k<4;k++)a[r][k]+=__shfl_xor_sync(0xffffffffu,a[r][k],o); }
      if(MDBC2){ v[0]+=__shfl_xor_sync(0xffffffffu,v[0],o); v[1]+=__shfl_xor_sync(0xffffffffu,v[1],o); v[2]+=__shfl_xor_sync(0xffffffffu,v[2],o); sub+=__shfl_xor_sync(0xffffffffu,sub,o); }
    }
  }
};

// The lanes of the warp stride over every neighbour row of g and first only TEST their candidate (one 128-bit position load, FP32
// distance).  About a quarter of the candidates of the window lie inside the kernel sphere, so the hits are compacted through a small
// per-warp queue in shared memory (ballot + prefix count) and the expensive part -- kernel, 20 to 25 FP64 accumulations -- always
// runs on 32 hits at once instead of on the few lanes of a row that happened to hit.
// sgn = +1: the matrix holds (r_g - r_j) (mDBC), -1: (r_j - r_g) (mDBC2), which is how the two reference kernels orient their
// first-order terms.
#define M_QLEN 64                // queue entries per warp: at most 31 waiting + 32 new
#define M_ROWS 32                // rows of the neighbour window: (2*scelldiv+1)^2 <= 25
struct GhostRows { uint32_t pbeg[M_ROWS], pmid[M_ROWS], excl[M_ROWS+1]; float sy[M_ROWS], sz[M_ROWS]; };   // per warp, shared memory

template<bool MDBC2> __device__ __forceinline__ void ghost_walk(const DevCte& c,const DivState* ds,const uint32_t* __restrict__ begt,
  const float4* __restrict__ posp,const float4* __restrict__ velrho,const typecode* __restrict__ code,bool fluidonly,const GhostNode& g,int lane,
  float4* qd,uint32_t* qi,GhostRows& rw,GhostSums<MDBC2>& u)
{
  const uint32_t* beg=begt+ds->box_fluid;
  const int ncx=int(ds->ncx), nsheet=int(ds->nsheet), cmx=int(ds->cellmin[0]);
  // the window touches at most two x blocks of the posp frame: A (of xini) and B (of xfin-1)
  const float sxa=g.gx+b2s_offx(c,g.acx,g.xini+cmx), sxb=g.gx+b2s_offx(c,g.acx,g.xfin-1+cmx);
  const int xmid=(((g.xfin-1+cmx)>>B2S_XBLK)<<B2S_XBLK)-cmx;             // first cell of block B
  const float sgn=(MDBC2? -1.f: 1.f);
  const uint32_t lt=(1u<<lane)-1u;
  // ---- all rows of the window at once: lane r reads the begin table of row r (one load latency instead of one per row), a warp
  //      scan of the row lengths turns the window into ONE dense list of candidates
  const int ny=g.yfin-g.yini, nrows=ny*(g.zfin-g.zini);
  uint32_t total;
  {
    uint32_t pb=0,pe=0,pm=0; float sy=0.f,sz=0.f;
    if(lane<nrows){
      const int y=g.yini+lane%ny, z=g.zini+lane/ny;
      const uint32_t* brow=beg+(y*ncx+z*nsheet);
      pb=brow[g.xini]; pe=brow[g.xfin];
      pm=(xmid<=g.xini? pb: (xmid>=g.xfin? pe: brow[xmid]));
      sy=g.gy+c.scell*float(g.cy-y); sz=g.gz+c.scell*float(g.cz-z);
    }
    const uint32_t len=pe-pb;
    uint32_t incl=len;
    #pragma unroll
    for(int o=1;o<32;o<<=1){ const uint32_t t=__shfl_up_sync(0xffffffffu,incl,o); if(lane>=o)incl+=t; }
    total=__shfl_sync(0xffffffffu,incl,31);
    __syncwarp();
    rw.pbeg[lane]=pb; rw.pmid[lane]=pm; rw.excl[lane]=incl-len; rw.sy[lane]=sy; rw.sz[lane]=sz;
    if(lane==31)rw.excl[32]=total;
    __syncwarp();
  }
  uint32_t q=0;                                                          // entries waiting in the queue (warp-uniform)
  // one hit of this lane: (W, gradW) and the sums
  auto hit_body=[&](const float4 e,uint32_t p2){
    if(fluidonly && !B2S_CODE_ISFLUID(code[p2]))return;
    const float drx=e.x, dry=e.y, drz=e.z;
    float fac;
    const float wab=wab_fac(c.tkernel,e.w,c,fac);
    const float4 v2=__ldg(velrho+p2);
    const float vol=c.massf/v2.w;
    const float w[4]={wab,fac*drx,fac*dry,fac*drz};                      // (W, gradW)
    const float d[4]={1.f,sgn*drx,sgn*dry,sgn*drz};
    #pragma unroll
    for(int r=0;r<4;r++){
      u.s[r]+=double(c.massf*w[r]);
      const float vw=w[r]*vol;
      #pragma unroll
      for(int k=0;k<4;k++)u.a[r][k]+=double(k? d[k]*vw: vw);
    }
    if(MDBC2){
      const float vw=wab*vol;
      u.v[0]+=double(vw*v2.x); u.v[1]+=double(vw*v2.y); u.v[2]+=double(vw*v2.z);
      u.sub-=double(vol*(drx*w[1]+dry*w[2]+drz*w[3]));
    }
  };
  // the first cnt (<=32) entries are evaluated, one per lane; the rest moves to the front
  auto drain=[&](uint32_t cnt){
    __syncwarp();
    if(uint32_t(lane)<cnt)hit_body(qd[lane],qi[lane]);
    const uint32_t rest=q-cnt;
    float4 e=make_float4(0,0,0,0); uint32_t i=0;
    if(uint32_t(lane)<rest){ e=qd[cnt+lane]; i=qi[cnt+lane]; }
    __syncwarp();
    if(uint32_t(lane)<rest){ qd[lane]=e; qi[lane]=i; }
    q=rest;
  };
  // ---- the dense candidate list, 32 at a time; a lane's row only ever moves forward
  int r=0;
  for(uint32_t c0=0;c0<total;c0+=32u){
    const uint32_t cnd=c0+uint32_t(lane);
    bool hit=false; float4 e=make_float4(0,0,0,0); uint32_t p2=0;
    if(cnd<total){
      while(cnd>=rw.excl[r+1])r++;
      p2=rw.pbeg[r]+(cnd-rw.excl[r]);
      const float4 pp2=__ldg(posp+p2);
      e.x=(p2>=rw.pmid[r]? sxb: sxa)-pp2.x; e.y=rw.sy[r]-pp2.y; e.z=rw.sz[r]-pp2.z;
      e.w=e.x*e.x+e.y*e.y+e.z*e.z;
      hit=(e.w<=c.kernelsize2);
    }
    const uint32_t m=__ballot_sync(0xffffffffu,hit);
    if(hit){ const uint32_t slot=q+__popc(m&lt); qd[slot]=e; qi[slot]=p2; }
    q+=__popc(m);
    if(q>=32u)drain(32u);
  }
  if(q)drain(q);
}

// Gaussian elimination with row pivoting on the augmented rows m[4][4+NB], everything in registers (all loops unroll, the row
// exchange is a predicated swap).  Leaves the NB solutions in the last NB columns and returns det(A) = signed product of pivots.
template<int NB> __device__ __forceinline__ double eliminate4(double (&m)[4][4+NB]){
  double det=1.0;
  #pragma unroll
  for(int k=0;k<4;k++){
    #pragma unroll
    for(int r=k+1;r<4;r++){
      const bool sw=(fabs(m[r][k])>fabs(m[k][k]));
      #pragma unroll
      for(int q=k;q<4+NB;q++){ const double t=m[k][q]; m[k][q]=(sw? m[r][q]: t); m[r][q]=(sw? t: m[r][q]); }
      det=(sw? -det: det);
    }
    det*=m[k][k];
    const double ipiv=1.0/m[k][k];
    #pragma unroll
    for(int r=k+1;r<4;r++){
      const double f=m[r][k]*ipiv;
      #pragma unroll
      for(int q=k+1;q<4+NB;q++)m[r][q]-=f*m[k][q];
    }
    #pragma unroll
    for(int q=k+1;q<4+NB;q++)m[k][q]*=ipiv;                               // row k normalised: back substitution needs no division
  }
  #pragma unroll
  for(int b=0;b<NB;b++){
    #pragma unroll
    for(int r=2;r>=0;r--){
      #pragma unroll
      for(int q=r+1;q<4;q++)m[r][4+b]-=m[r][q]*m[q][4+b];
    }
  }
  return det;
}

// ---- SlipMode 1 ("vel0"): density only ------------------------------------------------------------------------------------
// First-order result if |det A| >= 1e-3, Shepard if A00 > 0, else rho0 (JSphCpu_mdbc.cpp:112-147).
__global__ void __launch_bounds__(M_BS) k_mdbc_vel0(const DivState* __restrict__ ds, const uint32_t* __restrict__ begt,
  const double2* __restrict__ posxy, const double* __restrict__ posz, const typecode* __restrict__ code, const float4* __restrict__ boundnor,
  float4* posp, float4* velrho, const __grid_constant__ DevCte c)
{
  const uint32_t p1=blockIdx.x*M_WARPS+(threadIdx.x>>5);
  const int lane=threadIdx.x&31;
  const uint32_t n=(c.use_floating? ds->np: ds->npbok);    // floating particles may carry normals too (UseNormalsFt)
  if(p1>=n)return;
  const float4 nor=boundnor[p1];
  if(nor.x==0.f && nor.y==0.f && nor.z==0.f)return;
  const double2 xy=posxy[p1];
  const GhostNode g=ghost_node(c,ds,xy.x,xy.y,posz[p1],nor);
  __shared__ float4 s_qd[M_WARPS][M_QLEN];
  __shared__ uint32_t s_qi[M_WARPS][M_QLEN];
  __shared__ GhostRows s_rw[M_WARPS];
  GhostSums<false> u; u.zero();
  ghost_walk<false>(c,ds,begt,posp,velrho,code,c.use_floating!=0,g,lane,s_qd[threadIdx.x>>5],s_qi[threadIdx.x>>5],s_rw[threadIdx.x>>5],u);
  u.warp_total();
  const float b0=float(u.s[0]), b1=float(u.s[1]), b2=float(u.s[2]), b3=float(u.s[3]);
  double m[4][5];
  #pragma unroll
  for(int r=0;r<4;r++){
    #pragma unroll
    for(int k=0;k<4;k++)m[r][k]=u.a[r][k];
  }
  m[0][4]=double(b0); m[1][4]=double(b1); m[2][4]=double(b2); m[3][4]=double(b3);
  const double a00=u.a[0][0];
  const double det=eliminate4<1>(m);
  float rhofinal=c.rhopzero;
  if(fabs(det)>=double(1e-3f)){
    // value at the boundary particle = value at the ghost node + gradient . (particle - node), particle - node = -normal
    const float rhog=float(m[0][4]), gx=-float(m[1][4]), gy=-float(m[2][4]), gz=-float(m[3][4]);
    rhofinal=rhog+gx*(-nor.x)+gy*(-nor.y)+gz*(-nor.z);
  }
  else if(a00>0)rhofinal=float(double(b0)/a00);
  if(lane==0){
    float4 vr=velrho[p1]; vr.w=rhofinal; velrho[p1]=vr;
    float4 pp=posp[p1]; pp.w=b2s_press(rhofinal,c); posp[p1]=pp;
  }
}

// ---- SlipMode 2 (no-slip) / 3 (free-slip): mDBC2 ---------------------------------------------------------------------------
// JSphCpu_mdbc.cpp:246-436 with MDBC2_KEEPVEL as the reference defines it (DualSphDef.h:471): the boundary particle keeps its
// velocity, only its density is replaced.
//   submerged ghost node : boundmode=MDBC2; density from the first-order system if |det|>=1e-3 and the scaled condition number
//                          dp^2*|A|inf*|A^-1|inf <= 50 (JSphCpu_mdbc.cpp:222-229), else Shepard (sum V W < 0.1: max(rho0,Shepard));
//                          then pressure cloning P = cs0^2 (rho_g - rho0) + rho0 ((g - a_motion).n)(dpos.n), rho = rho0 + P/cs0^2
//                          (JSphCpu_mdbc.cpp:182-207); ghost velocity by Shepard sum; tangenvel = tangential part of
//                          (2 v_motion - v_ghost) [no-slip] or of v_ghost [free-slip]
//   not submerged        : boundmode=MDBC2OFF (the interaction then uses mass 0), rho=rho0, tangenvel from the motion velocity.
// v_motion = the velocity the boundary particle carries, a_motion = MotionAce of the last motion step (zero for fixed boundaries).
__global__ void __launch_bounds__(M_BS) k_mdbc2(const DivState* __restrict__ ds, const uint32_t* __restrict__ begt,
  const double2* __restrict__ posxy, const double* __restrict__ posz, const typecode* __restrict__ code, const float4* __restrict__ boundnor,
  float4* posp, float4* velrho, uint8_t* boundmode, float4* tangenvel, const float4* __restrict__ motionace, int freeslip, const __grid_constant__ DevCte c)
{
  const uint32_t p1=blockIdx.x*M_WARPS+(threadIdx.x>>5);
  const int lane=threadIdx.x&31;
  if(p1>=ds->npb)return;
  const float4 nor=boundnor[p1];
  if(nor.x==0.f && nor.y==0.f && nor.z==0.f){ if(lane==0)boundmode[p1]=B2S_BMODE_DBC; return; }
  const double2 xy=posxy[p1];
  const GhostNode g=ghost_node(c,ds,xy.x,xy.y,posz[p1],nor);
  __shared__ float4 s_qd[M_WARPS][M_QLEN];
  __shared__ uint32_t s_qi[M_WARPS][M_QLEN];
  __shared__ GhostRows s_rw[M_WARPS];
  GhostSums<true> u; u.zero();
  ghost_walk<true>(c,ds,begt,posp,velrho,code,true,g,lane,s_qd[threadIdx.x>>5],s_qi[threadIdx.x>>5],s_rw[threadIdx.x>>5],u);
  u.warp_total();
  const float snormal=sqrtf(nor.x*nor.x+nor.y*nor.y+nor.z*nor.z);
  const float nx=nor.x/snormal, ny=nor.y/snormal, nz=nor.z/snormal;
  const float4 vm=velrho[p1];
  float rhofinal, v2x=vm.x, v2y=vm.y, v2z=vm.z;
  uint8_t mode=B2S_BMODE_MDBC2OFF;
  if(float(u.sub)>0.f){
    mode=B2S_BMODE_MDBC2;
    const float b0=float(u.s[0]), b1=float(u.s[1]), b2=float(u.s[2]), b3=float(u.s[3]);
    const double a00=u.a[0][0];
    const float shepard=float(double(b0)/a00);
    if(float(a00)<0.1f)rhofinal=fmaxf(c.rhopzero,shepard);
    else{
      double m[4][8];
      float na=0.f;
      #pragma unroll
      for(int r=0;r<4;r++){
        double rowsum=0;
        #pragma unroll
        for(int k=0;k<4;k++){ m[r][k]=u.a[r][k]; m[r][4+k]=(r==k? 1.0: 0.0); rowsum+=fabs(u.a[r][k]); }
        na=fmaxf(na,float(rowsum));
      }
      const double det=eliminate4<4>(m);
      rhofinal=shepard;
      if(fabs(det)>=0.001){
        float ni=0.f;
        #pragma unroll
        for(int r=0;r<4;r++)ni=fmaxf(ni,float(fabs(m[r][4])+fabs(m[r][5])+fabs(m[r][6])+fabs(m[r][7])));
        if(c.dp2*na*ni<=50.f)rhofinal=float(m[0][4]*b0+m[0][5]*b1+m[0][6]*b2+m[0][7]*b3);
      }
    }
    const float pghost=float(c.cs0_d*c.cs0_d*double(rhofinal-c.rhopzero));
    const float normpos=(-nor.x)*nx+(-nor.y)*ny+(-nor.z)*nz;             // (particle - node) . n
    const float4 mace=(motionace? motionace[p1]: make_float4(0.f,0.f,0.f,0.f));
    const float normforce=c.rhopzero*((c.gravity[0]-mace.x)*nx+(c.gravity[1]-mace.y)*ny+(c.gravity[2]-mace.z)*nz);
    rhofinal=c.rhopzero+float(double(pghost+normforce*normpos)/(c.cs0_d*c.cs0_d));
    const float gvx=float(u.v[0]/a00), gvy=float(u.v[1]/a00), gvz=float(u.v[2]/a00);
    v2x=(freeslip? gvx: vm.x+vm.x-gvx); v2y=(freeslip? gvy: vm.y+vm.y-gvy); v2z=(freeslip? gvz: vm.z+vm.z-gvz);
  }
  else rhofinal=c.rhopzero;
  if(lane==0){
    const float vdn=v2x*nx+v2y*ny+v2z*nz;
    boundmode[p1]=mode;
    float4 vr=vm; vr.w=rhofinal; velrho[p1]=vr;
    float4 pp=posp[p1]; pp.w=b2s_press(rhofinal,c); posp[p1]=pp;
    tangenvel[p1]=make_float4(v2x-vdn*nx,v2y-vdn*ny,v2z-vdn*nz,0.f);
  }
}

int b2s_mdbc_impl(b2s_ctx* ctx){
  if(!ctx->boundnor)return b2s_fail(ctx,B2S_ERR_STATE,"mDBC needs boundary normals");
  if(ctx->prm.slipmode>=B2S_SLIP_NOSLIP){
    if(ctx->cte.use_floating)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"mDBC2 (SlipMode 2/3) with floating bodies is not supported");
    if(ctx->prm.casenmoving && ctx->slab_on)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"mDBC2 (SlipMode 2/3) with moving boundaries is not supported with slab decomposition");
    if(!ctx->npb_host)return B2S_OK;
    k_mdbc2<<<b2s_blocks(ctx->npb_host,M_WARPS),M_BS,0,ctx->stream>>>(ctx->div,(const uint32_t*)ctx->begincell,ctx->cur.posxy,ctx->cur.posz,ctx->cur.code,
      ctx->boundnor,ctx->posp,ctx->cur.velrho,ctx->boundmode,ctx->tangenvel,ctx->motionace,ctx->prm.slipmode==B2S_SLIP_FREESLIP? 1: 0,ctx->cte); B2S_LAUNCHED(ctx);
    B2S_CUDA(ctx,cudaGetLastError());
    return B2S_OK;
  }
  const uint32_t n=(ctx->cte.use_floating? ctx->np_host: ctx->npb_host);
  if(!n)return B2S_OK;
  k_mdbc_vel0<<<b2s_blocks(n,M_WARPS),M_BS,0,ctx->stream>>>(ctx->div,(const uint32_t*)ctx->begincell,ctx->cur.posxy,ctx->cur.posz,ctx->cur.code,
    ctx->boundnor,ctx->posp,ctx->cur.velrho,ctx->cte); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
