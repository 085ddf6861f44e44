// gauges.cu -- measurement gauges evaluated on the device every step: velocity at a point and surface water level along a line.
//
// Replaces JGaugeVelocity::CalculeCpuT / JGaugeSwl::CalculeCpuT (src/source/JDsGaugeItem.cpp:427-478, 728-790; GPU versions
// JDsGauge_ker.cu:50-310) as driven by JSphCpuSingle::RunGaugeSystem (JSphCpuSingle.cpp:1118-1125), which the reference calls
// right after ComputeStep and BEFORE the end-of-step cell divide: the particles already carry the new positions while the cell
// table is the one of the last divide.  The same state is used here (double positions, current begin table).
//   velocity: sum over fluid neighbours of W(r) m/rho * v (no kernel correction)
//   SWL     : mass(pt)=sum W(r) m/rho * m at PointNp+1 points from point0 towards point2; the surface is where the mass
//             falls below MassLimit, linearly interpolated between the two points around the crossing.
// Timing (computedt/outputdt) is the host's business (csrc/host/sim.cpp): the device evaluates all gauges after every step
// (a few microseconds) so that no host round trip is needed to know the time first.
#include "b2s_internal.h"
#include <float.h>

#define G_BS 128
enum { GAUGE_VEL=1, GAUGE_SWL=2, GAUGE_FORCE=3, GAUGE_MAXZ=4 };
struct GaugeDev { int type; uint32_t np, first; float masslimit; double p0[3], dir[3];
  uint32_t fcode, fidbegin, fcount, foff; };   // force gauge: type+value bits of the block, its id range, offset in the partace buffer

// fluid neighbours of a point: returns sum W m/rho (and the velocity sums)
__device__ void gauge_point(const DivState* ds,const uint32_t* begt,const double2* posxy,const double* posz,const typecode* code,
  const float4* velrho,const DevCte& c,double px,double py,double pz,double& sumwab,double& svx,double& svy,double& svz)
{
  sumwab=svx=svy=svz=0;
  // nsearch::Init (JCellSearch_inline.h): cell of the point relative to the cell domain, window clipped to the domain
  const int cx=int((px-c.domposmin[0])/c.scell_d)-int(ds->cellmin[0]);
  const int cy=int((py-c.domposmin[1])/c.scell_d)-int(ds->cellmin[1]);
  const int cz=int((pz-c.domposmin[2])/c.scell_d)-int(ds->cellmin[2]);
  const int d=c.scelldiv, ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const int xini=cx-min(cx,d), xfin=cx+min(ncx-cx-1,d)+1;
  const int yini=cy-min(cy,d), yfin=cy+min(ncy-cy-1,d)+1;
  const int zini=cz-min(cz,d), zfin=cz+min(ncz-cz-1,d)+1;
  if(xini<0 || yini<0 || zini<0 || xfin<=xini)return;
  const uint32_t* beg=begt+ds->box_fluid;
  for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++){
    const uint32_t pini=beg[y*ncx+z*nsheet+xini], pfin=beg[y*ncx+z*nsheet+xfin];
    for(uint32_t p2=pini;p2<pfin;p2++){
      const double2 xy=posxy[p2];
      const float drx=float(px-xy.x), dry=float(py-xy.y), drz=float(pz-posz[p2]);
      const float rr2=drx*drx+dry*dry+drz*drz;
      if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO && B2S_CODE_ISFLUID(code[p2])){
        float fac;
        float wab=wab_fac(c.tkernel,rr2,c,fac);
        const float4 v=velrho[p2];
        wab*=c.massf/v.w;
        sumwab+=wab; svx+=wab*v.x; svy+=wab*v.y; svz+=wab*v.z;
      }
    }
  }
}

// one thread per measurement point (velocity gauges have one, SWL gauges PointNp+1)
__global__ void __launch_bounds__(G_BS) k_gauge_points(const GaugeDev* __restrict__ gs, uint32_t ng, uint32_t npoints, const DivState* ds,
  const uint32_t* __restrict__ begt, const double2* __restrict__ posxy, const double* __restrict__ posz, const typecode* __restrict__ code,
  const float4* __restrict__ velrho, float* mass, float* res, const __grid_constant__ DevCte c, uint32_t own_cxa, uint32_t own_cxb)
{
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=npoints)return;
  uint32_t cg=0;
  while(cg+1<ng && gs[cg+1].first<=i)cg++;
  const GaugeDev& gd=gs[cg];
  const uint32_t cp=i-gd.first;
  const double px=gd.p0[0]+gd.dir[0]*double(cp), py=gd.p0[1]+gd.dir[1]*double(cp), pz=gd.p0[2]+gd.dir[2]*double(cp);
  double sw=0,vx=0,vy=0,vz=0;
  // slab runs: the rank whose slab holds the point's cell evaluates it (it has every neighbour, owned or ghost), the others add zeros
  const double acx=floor((px-c.domposmin[0])/c.scell_d);
  if(acx>=double(own_cxa) && acx<double(own_cxb))gauge_point(ds,begt,posxy,posz,code,velrho,c,px,py,pz,sw,vx,vy,vz);
  if(gd.type==GAUGE_VEL){ res[cg*4]=float(vx); res[cg*4+1]=float(vy); res[cg*4+2]=float(vz); res[cg*4+3]=0.f; }
  else mass[i]=float(sw*double(c.massf));
}
// SWL: the sequential search of JGaugeSwl::CalculeCpuT (JDsGaugeItem.cpp:770-782), one thread per gauge
__global__ void k_gauge_swl(const GaugeDev* __restrict__ gs, uint32_t ng, const float* __restrict__ mass, float* res){
  const uint32_t cg=blockIdx.x*blockDim.x+threadIdx.x;
  if(cg>=ng || gs[cg].type!=GAUGE_SWL)return;
  const GaugeDev& gd=gs[cg];
  const uint32_t pointnp=gd.np-1;
  double sx=DBL_MAX,sy=0,sz=0;
  float mpre=0;
  double px=gd.p0[0],py=gd.p0[1],pz=gd.p0[2];
  for(uint32_t cp=0;cp<=pointnp;cp++){
    const float m=mass[gd.first+cp];
    if(m>gd.masslimit)mpre=m;
    if(m<gd.masslimit && mpre){
      const float fxm1=(gd.masslimit-mpre)/(m-mpre)-1;
      sx=px+gd.dir[0]*double(fxm1); sy=py+gd.dir[1]*double(fxm1); sz=pz+gd.dir[2]*double(fxm1);
      break;
    }
    px+=gd.dir[0]; py+=gd.dir[1]; pz+=gd.dir[2];
  }
  if(sx==DBL_MAX){ const double k=(mpre? double(pointnp): 0.0); sx=gd.p0[0]+gd.dir[0]*k; sy=gd.p0[1]+gd.dir[1]*k; sz=gd.p0[2]+gd.dir[2]*k; }
  res[cg*4]=float(sx); res[cg*4+1]=float(sy); res[cg*4+2]=float(sz); res[cg*4+3]=0.f;
}

// ---- force gauge: JGaugeForce::CalculeCpuT (JDsGaugeItem.cpp:1847-1908; GPU JDsGauge_ker.cu:497-587) ----------------------
// Pressure force of the fluid on the boundary particles of one mk block: for every particle of the block the momentum-equation
// sum over its fluid neighbours (no viscous part, Wendland gradient also for cubic runs like the reference), stored by id,
// then one ordered sum times MassBound.  One thread per boundary particle, then one block per gauge.
__global__ void __launch_bounds__(G_BS) k_gauge_force_ace(GaugeDev gd, const DivState* ds, const uint32_t* __restrict__ begt, const double2* __restrict__ posxy,
  const double* __restrict__ posz, const typecode* __restrict__ code, const uint32_t* __restrict__ idp, const float4* __restrict__ velrho,
  float* partace, const __grid_constant__ DevCte c)
{
  const uint32_t p1=blockIdx.x*blockDim.x+threadIdx.x;
  if(p1>=ds->npbok)return;
  const typecode cod=code[p1];
  if(uint32_t(cod&(B2S_CODE_MASKTYPE|B2S_CODE_MASKVALUE))!=gd.fcode || !B2S_CODE_ISNORMAL(cod))return;   // periodic copies are ignored
  const uint32_t id=idp[p1];
  if(id<gd.fidbegin || id>=gd.fidbegin+gd.fcount)return;
  const double2 xy1=posxy[p1]; const double z1=posz[p1];
  const float4 v1=velrho[p1];
  const float press1=b2s_press(v1.w,c);
  float ax=0,ay=0,az=0;
  const int cx=int((xy1.x-c.domposmin[0])/c.scell_d)-int(ds->cellmin[0]);
  const int cy=int((xy1.y-c.domposmin[1])/c.scell_d)-int(ds->cellmin[1]);
  const int cz=int((z1   -c.domposmin[2])/c.scell_d)-int(ds->cellmin[2]);
  const int d=c.scelldiv, ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const int xini=cx-min(cx,d), xfin=cx+min(ncx-cx-1,d)+1;
  const int yini=cy-min(cy,d), yfin=cy+min(ncy-cy-1,d)+1;
  const int zini=cz-min(cz,d), zfin=cz+min(ncz-cz-1,d)+1;
  if(!(xini<0 || yini<0 || zini<0 || xfin<=xini)){
    const uint32_t* beg=begt+ds->box_fluid;
    for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++){
      const uint32_t pini=beg[y*ncx+z*nsheet+xini], pfin=beg[y*ncx+z*nsheet+xfin];
      for(uint32_t p2=pini;p2<pfin;p2++){
        const double2 xy=posxy[p2];
        const float drx=float(xy1.x-xy.x), dry=float(xy1.y-xy.y), drz=float(z1-posz[p2]);
        const float rr2=drx*drx+dry*dry+drz*drz;
        if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO && B2S_CODE_ISFLUID(code[p2])){
          float fac; wab_fac(B2S_KERNEL_WENDLAND,rr2,c,fac);
          const float4 v2=velrho[p2];
          const float press2=b2s_press(v2.w,c);
          const float prs=(press1+press2)/(v1.w*v2.w);
          const float p_vpm1=-prs*c.massf;
          ax+=p_vpm1*(fac*drx); ay+=p_vpm1*(fac*dry); az+=p_vpm1*(fac*drz);
        }
      }
    }
  }
  float* o=partace+size_t(gd.foff+(id-gd.fidbegin))*3;
  o[0]=ax; o[1]=ay; o[2]=az;
}
__global__ void __launch_bounds__(256) k_gauge_force_sum(GaugeDev gd, int cg, const float* __restrict__ partace, float* res, float massbound){
  __shared__ double sh[3][256];
  double s[3]={0,0,0};
  for(uint32_t i=threadIdx.x;i<gd.fcount;i+=256){ const float* a=partace+size_t(gd.foff+i)*3; s[0]+=a[0]; s[1]+=a[1]; s[2]+=a[2]; }
  for(int k=0;k<3;k++)sh[k][threadIdx.x]=s[k];
  __syncthreads();
  for(int o=128;o;o>>=1){ if(int(threadIdx.x)<o)for(int k=0;k<3;k++)sh[k][threadIdx.x]+=sh[k][threadIdx.x+o]; __syncthreads(); }
  if(threadIdx.x<3)res[cg*4+threadIdx.x]=float(sh[threadIdx.x][0])*massbound;
}

// ---- maxz gauge: JGaugeMaxZ::CalculeCpu (JDsGaugeItem.cpp:1023-1091; GPU JDsGauge_ker.cu:312-420) ------------------------
// Highest fluid particle within the horizontal distance DistLimit of the vertical line point0 .. point0+Height.  The reference
// walks the cell slabs from the top and stops at the first slab that holds such a particle; the maximum over all slabs of the
// window is the same number.  One block per gauge: threads stride over the (y,z) cell rows, then a block maximum.
// GaugeDev use: p0 = point0, dir[0] = Height, masslimit = DistLimit.
__global__ void __launch_bounds__(G_BS) k_gauge_maxz(GaugeDev gd, int cg, const DivState* ds, const uint32_t* __restrict__ begt, const double2* __restrict__ posxy,
  const double* __restrict__ posz, const typecode* __restrict__ code, float* res, const __grid_constant__ DevCte c, float* rawmax)
{
  __shared__ float sh[G_BS];
  const float distlimit=gd.masslimit, maxdist2=distlimit*distlimit;
  const int cx=int((gd.p0[0]-c.domposmin[0])/c.scell_d)-int(ds->cellmin[0]);
  const int cy=int((gd.p0[1]-c.domposmin[1])/c.scell_d)-int(ds->cellmin[1]);
  const int cz=int((gd.p0[2]-c.domposmin[2])/c.scell_d)-int(ds->cellmin[2]);
  const int cz2=int((gd.p0[2]+gd.dir[0]-c.domposmin[2])/c.scell_d)-int(ds->cellmin[2]);
  const int ncel=int(ceil(double(distlimit)/c.scell_d));   // Scell is the float cell size promoted to double in the reference too
  const int ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const int xini=max(cx-ncel,0), xfin=min(cx+ncel+1,ncx);
  const int yini=max(cy-ncel,0), yfin=min(cy+ncel+1,ncy);
  const int zini=max(cz-ncel,0), zfin=min(cz2+ncel+1,ncz);
  float zmax=-FLT_MAX;
  if(xini<xfin && yini<yfin && zini<zfin){
    const uint32_t* beg=begt+ds->box_fluid;
    const int ny=yfin-yini, nrows=ny*(zfin-zini);
    for(int r=int(threadIdx.x);r<nrows;r+=G_BS){
      const int y=yini+r%ny, z=zini+r/ny;
      const uint32_t pini=beg[y*ncx+z*nsheet+xini], pfin=beg[y*ncx+z*nsheet+xfin];
      for(uint32_t p2=pini;p2<pfin;p2++){
        const double pz=posz[p2];
        if(pz>double(zmax)){
          const double2 xy=posxy[p2];
          const float drx=float(gd.p0[0]-xy.x), dry=float(gd.p0[1]-xy.y);
          if(drx*drx+dry*dry<=maxdist2 && B2S_CODE_ISFLUID(code[p2]))zmax=float(pz);
        }
      }
    }
  }
  sh[threadIdx.x]=zmax;
  __syncthreads();
  for(int o=G_BS/2;o;o>>=1){ if(int(threadIdx.x)<o)sh[threadIdx.x]=fmaxf(sh[threadIdx.x],sh[threadIdx.x+o]); __syncthreads(); }
  if(threadIdx.x==0){
    if(rawmax){ rawmax[cg]=sh[0]; res[cg*4]=0.f; }   // slab runs: the maximum over the ranks comes first (k_gauge_maxz_finish)
    else res[cg*4]=fmaxf(sh[0],float(c.domposmin[2]));
    res[cg*4+1]=0.f; res[cg*4+2]=0.f; res[cg*4+3]=0.f;
  }
}
__global__ void k_gauge_maxz_finish(const GaugeDev* __restrict__ gs, uint32_t ng, const float* __restrict__ rawmax, float* res, float zmin){
  const uint32_t cg=blockIdx.x*blockDim.x+threadIdx.x;
  if(cg<ng && gs[cg].type==GAUGE_MAXZ)res[cg*4]=fmaxf(rawmax[cg],zmin);
}

static int gauges_upload(b2s_ctx* ctx, const std::vector<GaugeDev>& h){
  cudaFree(ctx->gauges_dev); cudaFree(ctx->gauge_res_dev); if(ctx->gauge_res_h)cudaFreeHost(ctx->gauge_res_h);   // gauge_mass_dev lives inside gauge_res_dev
  ctx->gauges_dev=NULL; ctx->gauge_res_dev=NULL; ctx->gauge_mass_dev=NULL; ctx->gauge_res_h=NULL;
  const size_t ng=h.size();
  B2S_CUDA(ctx,cudaMalloc(&ctx->gauges_dev,sizeof(GaugeDev)*ng));
  // one buffer [results 4*ng | point masses | raw maxima ng]: a slab run all-reduces the first two parts in one call (slab_comm.cu)
  const size_t npts=std::max<size_t>(ctx->gauge_points,1);
  B2S_CUDA(ctx,cudaMalloc((void**)&ctx->gauge_res_dev,sizeof(float)*(4*ng+npts+ng)));
  ctx->gauge_mass_dev=ctx->gauge_res_dev+4*ng;
  B2S_CUDA(ctx,cudaMallocHost((void**)&ctx->gauge_res_h,sizeof(float)*4*ng));
  B2S_CUDA(ctx,cudaMemcpy(ctx->gauges_dev,h.data(),sizeof(GaugeDev)*ng,cudaMemcpyHostToDevice));
  B2S_CUDA(ctx,cudaMemset(ctx->gauge_res_dev,0,sizeof(float)*(4*ng+npts+ng)));
  return B2S_OK;
}
// the host copy of the gauge list lives in the ctx as raw bytes (GaugeDev is private to this file)
static std::vector<GaugeDev> host_list(b2s_ctx* ctx){
  std::vector<GaugeDev> h(ctx->gauges_host.size()/sizeof(GaugeDev));
  if(!h.empty())memcpy(h.data(),ctx->gauges_host.data(),ctx->gauges_host.size());
  return h;
}
static void host_store(b2s_ctx* ctx,const std::vector<GaugeDev>& h){
  ctx->gauges_host.resize(h.size()*sizeof(GaugeDev));
  if(!h.empty())memcpy(ctx->gauges_host.data(),h.data(),ctx->gauges_host.size());
}

extern "C" int b2s_gauge_add_velocity(b2s_ctx* ctx, const double point[3], int* id){
  if(!ctx || !point)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  std::vector<GaugeDev> h=host_list(ctx);
  GaugeDev g; memset(&g,0,sizeof(g)); g.type=GAUGE_VEL; g.np=1; g.first=ctx->gauge_points;
  for(int k=0;k<3;k++)g.p0[k]=point[k];
  h.push_back(g); ctx->gauge_types.push_back(GAUGE_VEL); ctx->gauge_points+=1;
  host_store(ctx,h);
  if(id)*id=int(h.size())-1;
  return gauges_upload(ctx,h);
}
// Force gauge on the boundary block with ids [idbegin,idbegin+count) whose particles carry `typevalue` in the type+value bits
// of their code (fixed or moving; JGaugeSystem::AddGaugeForce, JDsGaugeSystem.cpp:472-495).
extern "C" int b2s_gauge_add_force(b2s_ctx* ctx, uint32_t idbegin, uint32_t count, uint32_t typevalue, int* id){
  if(!ctx || !count)return B2S_ERR_ARG;
  if((typevalue&B2S_CODE_MASKTYPE)>=B2S_CODE_TYPE_FLOATING || idbegin+count>ctx->prm.casenpb)
    return b2s_fail(ctx,B2S_ERR_ARG,"force gauge: only fixed or moving boundary blocks are allowed");
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  std::vector<GaugeDev> h=host_list(ctx);
  GaugeDev g; memset(&g,0,sizeof(g)); g.type=GAUGE_FORCE; g.np=0; g.first=ctx->gauge_points;
  g.fcode=typevalue&(B2S_CODE_MASKTYPE|B2S_CODE_MASKVALUE); g.fidbegin=idbegin; g.fcount=count; g.foff=ctx->gauge_force_n;
  ctx->gauge_force_n+=count;
  cudaFree(ctx->gauge_force_dev); ctx->gauge_force_dev=NULL;
  B2S_CUDA(ctx,cudaMalloc((void**)&ctx->gauge_force_dev,size_t(ctx->gauge_force_n)*12));
  h.push_back(g); ctx->gauge_types.push_back(GAUGE_FORCE);
  host_store(ctx,h);
  if(id)*id=int(h.size())-1;
  return gauges_upload(ctx,h);
}
// MaxZ gauge (JGaugeSystem::AddGaugeMaxZ): result[0] = zmax.
extern "C" int b2s_gauge_add_maxz(b2s_ctx* ctx, const double point0[3], double height, float distlimit, int* id){
  if(!ctx || !point0 || !(distlimit>0))return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  std::vector<GaugeDev> h=host_list(ctx);
  GaugeDev g; memset(&g,0,sizeof(g)); g.type=GAUGE_MAXZ; g.np=0; g.first=ctx->gauge_points; g.masslimit=distlimit;
  for(int k=0;k<3;k++)g.p0[k]=point0[k];
  g.dir[0]=height;
  h.push_back(g); ctx->gauge_types.push_back(GAUGE_MAXZ);
  host_store(ctx,h);
  if(id)*id=int(h.size())-1;
  return gauges_upload(ctx,h);
}
extern "C" int b2s_gauge_add_swl(b2s_ctx* ctx, const double point0[3], const double point2[3], double pointdp, float masslimit, int* id){
  if(!ctx || !point0 || !point2 || pointdp<=0 || masslimit<=0)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  std::vector<GaugeDev> h=host_list(ctx);
  // JGaugeSwl::SetPoints (JDsGaugeItem.cpp:614-633)
  double dv[3]={point2[0]-point0[0],point2[1]-point0[1],point2[2]-point0[2]};
  const double dis=sqrt(dv[0]*dv[0]+dv[1]*dv[1]+dv[2]*dv[2]);
  if(!(dis>0))return b2s_fail(ctx,B2S_ERR_ARG,"swl gauge: point0 and point2 coincide");
  unsigned pointnp=unsigned(dis/pointdp);
  if(dis-(pointdp*pointnp)>=pointdp*0.1)pointnp++;
  if(pointnp<1)pointnp++;
  const double dp=dis/pointnp;
  GaugeDev g; memset(&g,0,sizeof(g)); g.type=GAUGE_SWL; g.np=pointnp+1; g.first=ctx->gauge_points; g.masslimit=masslimit;
  for(int k=0;k<3;k++){ g.p0[k]=point0[k]; g.dir[k]=dv[k]/dis*dp; }
  h.push_back(g); ctx->gauge_types.push_back(GAUGE_SWL); ctx->gauge_points+=g.np;
  host_store(ctx,h);
  if(id)*id=int(h.size())-1;
  return gauges_upload(ctx,h);
}
// phase 0: everything (one context).  Slab runs (slab_comm.cu grp_gauges): phase 1 = the partial results of this rank (points it owns,
// boundary particles it owns, the highest particle it sees), then the all-reduce over the ranks (sum of results and point masses,
// maximum of the raw maxima), then phase 2 = what follows from the reduced values (surface search, maxz clamp) and the copy to the host.
int b2s_gauges_run(b2s_ctx* ctx, int phase){
  const uint32_t ng=uint32_t(ctx->gauge_types.size());
  if(!ng)return B2S_OK;
  cudaStream_t s=ctx->stream;
  const bool slab=(phase!=0);
  float* rawmax=ctx->gauge_mass_dev+std::max<size_t>(ctx->gauge_points,1);
  if(phase!=2){
    if(slab)B2S_CUDA(ctx,cudaMemsetAsync(ctx->gauge_res_dev,0,sizeof(float)*(4*size_t(ng)+std::max<size_t>(ctx->gauge_points,1)),s));
    const std::vector<GaugeDev> h=host_list(ctx);
    for(uint32_t cg=0;cg<ng;cg++)if(h[cg].type==GAUGE_MAXZ){
      k_gauge_maxz<<<1,G_BS,0,s>>>(h[cg],int(cg),ctx->div,(const uint32_t*)ctx->begincell,ctx->cur.posxy,ctx->cur.posz,ctx->cur.code,ctx->gauge_res_dev,ctx->cte,slab? rawmax: (float*)NULL); B2S_LAUNCHED(ctx);
    }
    if(ctx->gauge_force_n){
      B2S_CUDA(ctx,cudaMemsetAsync(ctx->gauge_force_dev,0,size_t(ctx->gauge_force_n)*12,s));
      for(uint32_t cg=0;cg<ng;cg++)if(h[cg].type==GAUGE_FORCE && ctx->npbok_host){
        k_gauge_force_ace<<<b2s_blocks(ctx->npbok_host,G_BS),G_BS,0,s>>>(h[cg],ctx->div,(const uint32_t*)ctx->begincell,ctx->cur.posxy,ctx->cur.posz,
          ctx->cur.code,ctx->cur.idp,ctx->cur.velrho,ctx->gauge_force_dev,ctx->cte); B2S_LAUNCHED(ctx);
        k_gauge_force_sum<<<1,256,0,s>>>(h[cg],int(cg),ctx->gauge_force_dev,ctx->gauge_res_dev,ctx->cte.massb); B2S_LAUNCHED(ctx);
      }
    }
    const uint32_t cxa=(slab? ctx->slab_cxa: 0u), cxb=(slab? ctx->slab_cxb: 0xffffffffu);
    if(ctx->gauge_points)k_gauge_points<<<b2s_blocks(ctx->gauge_points,G_BS),G_BS,0,s>>>((const GaugeDev*)ctx->gauges_dev,ng,ctx->gauge_points,ctx->div,(const uint32_t*)ctx->begincell,
      ctx->cur.posxy,ctx->cur.posz,ctx->cur.code,ctx->cur.velrho,ctx->gauge_mass_dev,ctx->gauge_res_dev,ctx->cte,cxa,cxb); B2S_LAUNCHED(ctx);
  }
  if(phase!=1){
    if(slab){ k_gauge_maxz_finish<<<b2s_blocks(ng,32),32,0,s>>>((const GaugeDev*)ctx->gauges_dev,ng,rawmax,ctx->gauge_res_dev,float(ctx->cte.domposmin[2])); B2S_LAUNCHED(ctx); }
    k_gauge_swl<<<b2s_blocks(ng,32),32,0,s>>>((const GaugeDev*)ctx->gauges_dev,ng,ctx->gauge_mass_dev,ctx->gauge_res_dev); B2S_LAUNCHED(ctx);
    B2S_CUDA(ctx,cudaMemcpyAsync(ctx->gauge_res_h,ctx->gauge_res_dev,sizeof(float)*4*ng,cudaMemcpyDeviceToHost,s));
  }
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
extern "C" int b2s_gauges_compute(b2s_ctx* ctx){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  if(ctx->slab_on && ctx->slab_nranks>1)return b2s_fail(ctx,B2S_ERR_STATE,"gauges of a slab rank are evaluated collectively: b2s_slab_gauges_compute");
  int r=b2s_gauges_run(ctx,0); if(r)return r;
  B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  return B2S_OK;
}
extern "C" int b2s_gauge_result(b2s_ctx* ctx, int id, float out[3]){
  if(!ctx || !out || id<0 || size_t(id)>=ctx->gauge_types.size() || !ctx->gauge_res_h)return B2S_ERR_ARG;
  for(int k=0;k<3;k++)out[k]=ctx->gauge_res_h[id*4+k];
  return B2S_OK;
}
