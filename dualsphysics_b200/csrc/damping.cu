// damping.cu -- damping zones (passive wave absorption): velocity reduction of the fluid particles inside a zone.
//
// Replaces (under /root/reference/src/source):
//   JDsDampingOp_Plane::ReadXml / ComputeDomPlanes / ComputeDampingCpu   JDsDamping.cpp:58-131,185-233  (GPU: cusph::ComputeDampingPlane(Dom), JSphGpu_ker.cu:2445-2540)
//   JDsDampingOp_Box::ReadXml / ComputeDampingCpu                        JDsDamping.cpp:345-391,446-492 (GPU: cusph::ComputeDampingBox, JSphGpu_ker.cu:2545-2625)
//   JDsDamping::ComputeDampingCpu, JSphCpu::RunDamping                   JDsDamping.cpp:713-720, JSphCpu.cpp:2312-2316
// called once per step after the integrator and the floating-body update (JSphCpuSingle.cpp:885,928).
// The reference launches one kernel per zone; here one kernel applies every zone in list order to its particle
// (same result: each zone only rescales the velocity it finds), reading dt from the device-resident StepState.
#include "b2s_internal.h"
#include <math.h>

struct DampOp {
  int32_t type;            // 0 plane (dampingzone), 1 box (dampingbox)
  int32_t usedomain;       // plane only
  float over, redumax, factor[3], dist;
  double plane[4];                      // normalised plane through limitmin, normal limitmin->limitmax
  double zmin, zmax, dompla[4][4];      // optional <domain>: z range + 4 lateral planes
  double limmin1[3], limmin2[3], limmax1[3], limmax2[3], limover1[3], limover2[3], boxsize1[3], boxsize2[3];
};

__device__ __forceinline__ double d_planepoint(const double* pl,double x,double y,double z){ return pl[0]*x+pl[1]*y+pl[2]*z+pl[3]; }
__device__ __forceinline__ bool d_le3(const double* a,double x,double y,double z){ return a[0]<=x && a[1]<=y && a[2]<=z; }
__device__ __forceinline__ bool d_ge3(const double* a,double x,double y,double z){ return x<=a[0] && y<=a[1] && z<=a[2]; }

__global__ void __launch_bounds__(256) k_damping(const DampOp* __restrict__ ops,int nops,const DivState* __restrict__ ds,const StepState* __restrict__ st,
  const typecode* __restrict__ code,const double2* __restrict__ posxy,const double* __restrict__ posz,float4* velrho)
{
  const uint32_t p=ds->npb+blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=ds->np)return;
  const typecode cod=code[p];
  if(!(B2S_CODE_ISNORMAL(cod) && B2S_CODE_ISFLUID(cod)))return;   // floating, slab ghosts and excluded particles are skipped
  const double2 xy=posxy[p]; const double x=xy.x, y=xy.y, z=posz[p];
  const double dt=st->dt;
  float4 v=velrho[p];
  bool changed=false;
  for(int k=0;k<nops;k++){
    const DampOp& o=ops[k];
    double fdis;
    if(o.type==0){
      const double vdis=d_planepoint(o.plane,x,y,z);
      if(!(0<vdis && vdis<=double(o.dist+o.over)))continue;
      if(o.usedomain && !(z>=o.zmin && z<=o.zmax && d_planepoint(o.dompla[0],x,y,z)<=0 && d_planepoint(o.dompla[1],x,y,z)<=0
        && d_planepoint(o.dompla[2],x,y,z)<=0 && d_planepoint(o.dompla[3],x,y,z)<=0))continue;
      fdis=(vdis>=double(o.dist)? 1.: vdis/double(o.dist));
    }
    else{
      if(!(d_le3(o.limover1,x,y,z) && d_ge3(o.limover2,x,y,z)))continue;   // outside the overlimit domain
      if(d_le3(o.limmin1,x,y,z) && d_ge3(o.limmin2,x,y,z))continue;         // inside the free domain
      fdis=1.;
      if(d_le3(o.limmax1,x,y,z) && d_ge3(o.limmax2,x,y,z)){
        fdis=0;
        if(o.boxsize2[2]){ const double f=(z-o.limmin2[2])/o.boxsize2[2]; fdis=(fdis>=f? fdis: f); }
        if(o.boxsize2[1]){ const double f=(y-o.limmin2[1])/o.boxsize2[1]; fdis=(fdis>=f? fdis: f); }
        if(o.boxsize2[0]){ const double f=(x-o.limmin2[0])/o.boxsize2[0]; fdis=(fdis>=f? fdis: f); }
        if(o.boxsize1[2]){ const double f=(o.limmin1[2]-z)/o.boxsize1[2]; fdis=(fdis>=f? fdis: f); }
        if(o.boxsize1[1]){ const double f=(o.limmin1[1]-y)/o.boxsize1[1]; fdis=(fdis>=f? fdis: f); }
        if(o.boxsize1[0]){ const double f=(o.limmin1[0]-x)/o.boxsize1[0]; fdis=(fdis>=f? fdis: f); }
      }
    }
    const double redudt=dt*(fdis*fdis)*double(o.redumax);
    double rx=1.-redudt*double(o.factor[0]), ry=1.-redudt*double(o.factor[1]), rz=1.-redudt*double(o.factor[2]);
    rx=(rx<0? 0.: rx); ry=(ry<0? 0.: ry); rz=(rz<0? 0.: rz);
    v.x=float(rx*double(v.x)); v.y=float(ry*double(v.y)); v.z=float(rz*double(v.z));
    changed=true;
  }
  if(changed)velrho[p]=v;
}

int b2s_damping_impl(b2s_ctx* ctx){
  if(ctx->damp_host.empty())return B2S_OK;
  const uint32_t npf=ctx->npf_host;
  if(!npf)return B2S_OK;
  const int nops=int(ctx->damp_host.size()/sizeof(DampOp));
  if(ctx->damp_dirty){
    if(ctx->damp_dev)cudaFree(ctx->damp_dev);
    B2S_CUDA(ctx,cudaMalloc(&ctx->damp_dev,ctx->damp_host.size()));
    B2S_CUDA(ctx,cudaMemcpyAsync(ctx->damp_dev,ctx->damp_host.data(),ctx->damp_host.size(),cudaMemcpyHostToDevice,ctx->stream));
    ctx->damp_dirty=false;
  }
  k_damping<<<b2s_blocks(npf,256),256,0,ctx->stream>>>((const DampOp*)ctx->damp_dev,nops,ctx->div,ctx->step,ctx->cur.code,ctx->cur.posxy,ctx->cur.posz,ctx->cur.velrho);
  B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}

static void push_op(b2s_ctx* ctx,const DampOp& o){
  const unsigned char* b=(const unsigned char*)&o;
  ctx->damp_host.insert(ctx->damp_host.end(),b,b+sizeof(DampOp));
  ctx->damp_dirty=true;
}

// fgeo::Plane3Pt (FunGeo3d.cpp:56-66)
static void plane3pt(const double p1[3],const double p2[3],const double p3[3],double pl[4]){
  pl[0]=pl[1]=pl[2]=pl[3]=0;
  auto same=[](const double* a,const double* b){ return a[0]==b[0] && a[1]==b[1] && a[2]==b[2]; };
  if(same(p1,p2) || same(p1,p3) || same(p2,p3))return;
  const double v1[3]={p2[0]-p1[0],p2[1]-p1[1],p2[2]-p1[2]}, v2[3]={p3[0]-p1[0],p3[1]-p1[1],p3[2]-p1[2]};
  const double v[3]={v1[1]*v2[2]-v1[2]*v2[1], v1[2]*v2[0]-v1[0]*v2[2], v1[0]*v2[1]-v1[1]*v2[0]};
  pl[0]=v[0]; pl[1]=v[1]; pl[2]=v[2]; pl[3]=-(v[0]*p1[0]+v[1]*p1[1]+v[2]*p1[2]);
}

extern "C" int b2s_add_damping_plane(b2s_ctx* ctx, const double limitmin[3], const double limitmax[3], float overlimit, float redumax,
  const float factorxyz[3], const double* domain)
{
  if(!ctx || !limitmin || !limitmax || !factorxyz)return B2S_ERR_ARG;
  DampOp o; memset(&o,0,sizeof(o));
  o.type=0; o.over=overlimit; o.redumax=redumax; for(int k=0;k<3;k++)o.factor[k]=factorxyz[k];
  // Dist=|limitmax-limitmin|, Plane=PlanePtVec(limitmin,vec) with the unit vector (JDsDamping.cpp:124-130, FunGeo3d.h:545-548)
  const double vec[3]={limitmax[0]-limitmin[0],limitmax[1]-limitmin[1],limitmax[2]-limitmin[2]};
  const double m=sqrt(vec[0]*vec[0]+vec[1]*vec[1]+vec[2]*vec[2]);
  if(!(m>0))return b2s_fail(ctx,B2S_ERR_ARG,"damping zone: limitmin and limitmax coincide");
  o.dist=float(m);
  for(int k=0;k<3;k++)o.plane[k]=vec[k]/m;
  o.plane[3]=-o.plane[0]*limitmin[0]-o.plane[1]*limitmin[1]-o.plane[2]*limitmin[2];
  if(domain){   // zmin, zmax, point1..4 (x,y): ComputeDomPlanes (JDsDamping.cpp:58-91)
    o.usedomain=1; o.zmin=domain[0]; o.zmax=domain[1];
    double vp[4][2]; for(int c=0;c<4;c++){ vp[c][0]=domain[2+2*c]; vp[c][1]=domain[3+2*c]; }
    double ang[4]={0,0,0,0};
    for(int c=1;c<4;c++){ ang[c]=atan2(vp[0][1]-vp[c][1],vp[0][0]-vp[c][0])*(180./3.14159265358979323846); if(ang[c]<0)ang[c]+=360.; }
    for(int c=1;c<3;c++)for(int c2=c+1;c2<4;c2++)if(ang[c]>ang[c2]){ std::swap(ang[c],ang[c2]); std::swap(vp[c][0],vp[c2][0]); std::swap(vp[c][1],vp[c2][1]); }
    for(int c=0;c<4;c++){
      const int n=(c+1)&3;
      const double a[3]={vp[c][0],vp[c][1],0}, b[3]={vp[n][0],vp[n][1],0}, t[3]={vp[n][0],vp[n][1],1};
      plane3pt(a,b,t,o.dompla[c]);
    }
  }
  push_op(ctx,o);
  return B2S_OK;
}

extern "C" int b2s_add_damping_box(b2s_ctx* ctx, const double limitmin_ini[3], const double limitmin_end[3], const double limitmax_ini[3], const double limitmax_end[3],
  float overlimit, float redumax, const float factorxyz[3], uint32_t directions)
{
  if(!ctx || !limitmin_ini || !limitmin_end || !limitmax_ini || !limitmax_end || !factorxyz)return B2S_ERR_ARG;
  DampOp o; memset(&o,0,sizeof(o));
  o.type=1; o.over=overlimit; o.redumax=redumax; for(int k=0;k<3;k++)o.factor[k]=factorxyz[k];
  // checks and derived boxes of JDsDampingOp_Box::ReadXml (JDsDamping.cpp:365-391)
  auto le=[](const double* a,const double* b){ return a[0]<=b[0] && a[1]<=b[1] && a[2]<=b[2]; };
  if(!le(limitmin_ini,limitmin_end))return b2s_fail(ctx,B2S_ERR_ARG,"damping box: begin of LimitMin is higher than end of LimitMin");
  if(!le(limitmax_ini,limitmax_end))return b2s_fail(ctx,B2S_ERR_ARG,"damping box: begin of LimitMax is higher than end of LimitMax");
  if(!le(limitmax_ini,limitmin_ini) || !le(limitmin_end,limitmax_end))return b2s_fail(ctx,B2S_ERR_ARG,"damping box: LimitMin box is not inside of LimitMax box");
  if(overlimit<0)return b2s_fail(ctx,B2S_ERR_ARG,"damping box: OverLimit must not be negative");
  if(!(directions&B2S_DAMPDIR_ALL) || (directions&~uint32_t(B2S_DAMPDIR_ALL)))return b2s_fail(ctx,B2S_ERR_ARG,"damping box: at least one valid direction must be configured");
  for(int k=0;k<3;k++){
    o.limmin1[k]=limitmin_ini[k]; o.limmin2[k]=limitmin_end[k]; o.limmax1[k]=limitmax_ini[k]; o.limmax2[k]=limitmax_end[k];
    o.limover1[k]=o.limmax1[k]-double(overlimit); o.limover2[k]=o.limmax2[k]+double(overlimit);
  }
  // axis k: low side flag, high side flag (left/right = x, front/back = y, bottom/top = z)
  const uint32_t lowf[3]={B2S_DAMPDIR_LEFT,B2S_DAMPDIR_FRONT,B2S_DAMPDIR_BOTTOM}, highf[3]={B2S_DAMPDIR_RIGHT,B2S_DAMPDIR_BACK,B2S_DAMPDIR_TOP};
  for(int k=0;k<3;k++){
    if(!(directions&highf[k]))o.limmax2[k]=o.limover2[k]=o.limmin2[k];
    if(!(directions&lowf[k])) o.limmax1[k]=o.limover1[k]=o.limmin1[k];
  }
  for(int k=0;k<3;k++){
    o.boxsize1[k]=o.limmin1[k]-o.limmax1[k]; o.boxsize2[k]=o.limmax2[k]-o.limmin2[k];
    if(!(directions&highf[k]) || o.boxsize2[k]<0)o.boxsize2[k]=0;
    if(!(directions&lowf[k])  || o.boxsize1[k]<0)o.boxsize1[k]=0;
  }
  push_op(ctx,o);
  return B2S_OK;
}
