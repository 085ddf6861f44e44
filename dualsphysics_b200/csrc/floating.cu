// floating.cu -- free-floating rigid bodies: force/torque reduction, rigid-body update.
//
// Replaces JSphGpuSingle::RunFloating (src/source/JSphGpuSingle.cpp:827-907):
//   cusph::FtPartsSumAce  (KerFtPartsSumAce, JSphGpu_ker.cu:1938-2004; CPU JSphCpuSingle.cpp:1014-1038)
//   JSph::FtComputeAceVel (JSph.cpp:2565-2653, host code in the reference)
//   cusph::FtPartsUpdate  (KerFtPartsUpdate, JSphGpu_ker.cu:2027-2074; CPU JSphCpuSingle.cpp:1043-1108)
//   JSph::FtUpdateFloatings (JSph.cpp:2687-2701)
// The reference sums each body in ONE block of 256 threads whose thread 0 adds the 256 partials serially, copies the
// sums to the host, does the rigid-body algebra on the host and copies the result back.  Here the reduction is a
// deterministic two-level tree spread over as many blocks as the body needs (k_ft_sum_blocks -> k_ft_compute), and the
// rigid-body algebra runs on the device (one thread per body), so RunFloating adds no host round trip to the step.
#include "b2s_internal.h"
#include <math.h>

#define FT_BS 256

struct FtDev {
  b2s_floating f;
  float nvel[3], nomega[3];      // fto_vellinang
  double ncenter[3];             // fto_center
  float sums[6];                 // fto_acelinang input: sum ace, sum r x ace
};

// ---- level 1: per-block partial sums (fixed tree order -> deterministic) ------------------------------------------
__global__ void __launch_bounds__(FT_BS) k_ft_sum_blocks(const FtDev* __restrict__ fts, uint32_t cf, const DivState* __restrict__ ds,
  const uint32_t* __restrict__ ridp, uint32_t casenfixed, const double2* __restrict__ posxy, const double* __restrict__ posz,
  const typecode* __restrict__ code, const float4* __restrict__ acear, float* __restrict__ partials)
{
  __shared__ float sh[6][FT_BS/32];
  const b2s_floating& fo=fts[cf].f;
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  float v[6]={0,0,0,0,0,0};
  if(i<fo.count){
    const uint32_t p=ridp[fo.begin-casenfixed+i];
    if(p<ds->np && B2S_CODE_ISNORMAL(code[p])){   // slab mode: halo ghosts belong to the neighbour's sum
      const float4 a=acear[p];
      const double2 xy=posxy[p];
      const float dx=float(xy.x-fo.center[0]), dy=float(xy.y-fo.center[1]), dz=float(posz[p]-fo.center[2]);
      v[0]=a.x; v[1]=a.y; v[2]=a.z;
      v[3]=a.z*dy-a.y*dz; v[4]=a.x*dz-a.z*dx; v[5]=a.y*dx-a.x*dy;
    }
  }
  #pragma unroll
  for(int k=0;k<6;k++){
    float x=v[k];
    #pragma unroll
    for(int o=16;o;o>>=1)x+=__shfl_down_sync(0xffffffffu,x,o);
    if((threadIdx.x&31)==0)sh[k][threadIdx.x>>5]=x;
  }
  __syncthreads();
  if(threadIdx.x<6){
    float s=0;
    #pragma unroll
    for(int w=0;w<FT_BS/32;w++)s+=sh[threadIdx.x][w];
    partials[size_t(blockIdx.x)*6+threadIdx.x]=s;
  }
}

__device__ __forceinline__ void ft_constraints(int con,float* lin,float* ang){
  if(con&1)lin[0]=0; if(con&2)lin[1]=0; if(con&4)lin[2]=0;
  if(con&8)ang[0]=0; if(con&16)ang[1]=0; if(con&32)ang[2]=0;
}

// ---- level 2 + JSph::FtComputeAceVel on the device: one block per call, thread 0..5 finish the sums --------------------
__global__ void k_ft_reduce(uint32_t cf, uint32_t nblocks, const float* __restrict__ partials, float* __restrict__ allsums){
  if(threadIdx.x<6){
    float s=0;
    for(uint32_t b=0;b<nblocks;b++)s+=partials[size_t(b)*6+threadIdx.x];
    allsums[size_t(cf)*6+threadIdx.x]=s;      // slab mode: all-reduced (SUM) over the ranks before k_ft_compute
  }
}
__global__ void k_ft_compute(FtDev* fts, uint32_t cf, const float* __restrict__ allsums, const StepState* __restrict__ st,
  int predictor, double ftpause, const __grid_constant__ DevCte c)
{
  if(threadIdx.x)return;
  float sums[6];
  for(int k=0;k<6;k++)sums[k]=allsums[size_t(cf)*6+k];
  FtDev& d=fts[cf];
  b2s_floating& fo=d.f;
  for(int k=0;k<6;k++)d.sums[k]=sums[k];
  const double dt=(predictor? st->dt*.5: st->dt);
  if(st->timestep<ftpause)return;
  const float fmassp=fo.massp, fmass=fo.mass;
  float flin[3],fang[3],acelin[3],aceang[3];
  for(int k=0;k<3;k++){ flin[k]=sums[k]*fmassp; fang[k]=sums[3+k]*fmassp; }
  for(int k=0;k<3;k++)acelin[k]=(flin[k]+(c.gravity[k]*fmass))/fmass;
  {
    // I = (R*I0)*R^T with the accumulated rotation (fmath::RotMatrix3x3, FunctionsMath.h:345-353), alpha = I^-1 * torque
    const float ax=fo.angles[0], ay=fo.angles[1], az=fo.angles[2];
    const float cx=cosf(ax),cy=cosf(ay),cz=cosf(az),sx=sinf(ax),sy=sinf(ay),sz=sinf(az);
    const float R[3][3]={{cy*cz,-cy*sz,sy},{sx*sy*cz+cx*sz,-sx*sy*sz+cx*cz,-sx*cy},{-cx*sy*cz+sx*sz,cx*sy*sz+sx*cz,cx*cy}};
    float A[3][3],I[3][3];
    for(int i=0;i<3;i++)for(int j=0;j<3;j++){ float s=0; for(int k=0;k<3;k++)s+=R[i][k]*fo.inertiaini[k*3+j]; A[i][j]=s; }
    for(int i=0;i<3;i++)for(int j=0;j<3;j++){ float s=0; for(int k=0;k<3;k++)s+=A[i][k]*R[j][k]; I[i][j]=s; }
    const float det=I[0][0]*I[1][1]*I[2][2]+I[0][1]*I[1][2]*I[2][0]+I[0][2]*I[1][0]*I[2][1]-I[2][0]*I[1][1]*I[0][2]-I[2][1]*I[1][2]*I[0][0]-I[2][2]*I[1][0]*I[0][1];
    float inv[3][3]={{0,0,0},{0,0,0},{0,0,0}};
    if(det){
      inv[0][0]= (I[1][1]*I[2][2]-I[1][2]*I[2][1])/det; inv[0][1]=-(I[0][1]*I[2][2]-I[0][2]*I[2][1])/det; inv[0][2]= (I[0][1]*I[1][2]-I[0][2]*I[1][1])/det;
      inv[1][0]=-(I[1][0]*I[2][2]-I[1][2]*I[2][0])/det; inv[1][1]= (I[0][0]*I[2][2]-I[0][2]*I[2][0])/det; inv[1][2]=-(I[0][0]*I[1][2]-I[0][2]*I[1][0])/det;
      inv[2][0]= (I[1][0]*I[2][1]-I[1][1]*I[2][0])/det; inv[2][1]=-(I[0][0]*I[2][1]-I[0][1]*I[2][0])/det; inv[2][2]= (I[0][0]*I[1][1]-I[0][1]*I[1][0])/det;
    }
    for(int i=0;i<3;i++)aceang[i]=fang[0]*inv[i][0]+fang[1]*inv[i][1]+fang[2]*inv[i][2];
  }
  if(!predictor)for(int k=0;k<3;k++){ fo.fluforcelin[k]=flin[k]; fo.fluforceang[k]=fang[k]; fo.preacelin[k]=acelin[k]; fo.preaceang[k]=aceang[k]; }   // saveftvalues (JSph.cpp:2613-2620)
  float fvel[3],fomega[3];
  for(int k=0;k<3;k++){
    d.ncenter[k]=fo.center[k]+dt*fo.fvel[k];
    fvel[k]=float(dt*acelin[k]+fo.fvel[k]);
    fomega[k]=float(dt*aceang[k]+fo.fomega[k]);
  }
  if(fo.constraints){ ft_constraints(fo.constraints,acelin,aceang); ft_constraints(fo.constraints,fvel,fomega); }
  for(int k=0;k<3;k++){ d.nvel[k]=fvel[k]; d.nomega[k]=fomega[k]; }
  if(!predictor){
    // JSph::FtUpdateFloatings (JSph.cpp:2687-2701); the particle update below reads nvel/nomega/ncenter
    for(int k=0;k<3;k++){
      fo.angles[k]=float(double(fo.angles[k])+double(fomega[k])*dt);
      fo.facelin[k]=(fvel[k]-fo.fvel[k])/float(dt);
      fo.faceang[k]=(fomega[k]-fo.fomega[k])/float(dt);
    }
  }
}

// ---- KerFtPartsUpdate: one thread per floating particle -------------------------------------------------------------------------
__global__ void __launch_bounds__(FT_BS) k_ft_update(const FtDev* __restrict__ fts, uint32_t cf, const DivState* __restrict__ ds, const StepState* __restrict__ st,
  int predictor, double ftpause, const uint32_t* __restrict__ ridp, uint32_t casenfixed, double2* posxy, double* posz, float4* velrho,
  uint32_t* dcell, typecode* code, const __grid_constant__ DevCte c)
{
  const FtDev& d=fts[cf];
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=d.f.count || st->timestep<ftpause)return;
  const uint32_t p=ridp[d.f.begin-casenfixed+i];
  if(p>=ds->np || !B2S_CODE_ISNORMAL(code[p]))return;   // not on this rank, or a halo ghost (frozen)
  const double dt=(predictor? st->dt*.5: st->dt);
  float4 vr=velrho[p];
  double2 xy=posxy[p]; double z=posz[p];
  const double mx=dt*double(vr.x), my=dt*double(vr.y), mz=dt*double(vr.z);
  // JSphCpu::UpdatePos with outrho=false (JSphCpu.cpp:1458-1510)
  const bool outmov=(fabsf(float(mx))>c.movlimit || fabsf(float(my))>c.movlimit || fabsf(float(mz))>c.movlimit);
  xy.x+=mx; xy.y+=my; z+=mz;
  uint32_t cell;
  const bool out=b2s_pos_to_cell(c,xy.x,xy.y,z,cell);
  posxy[p]=xy; posz[p]=z;
  if(outmov || out){
    code[p]=typecode(B2S_CODE_SETNORMAL(code[p])|(out? B2S_CODE_OUTPOS: B2S_CODE_OUTMOV));
    dcell[p]=B2S_DCEL_MAPOUT;
  }
  else dcell[p]=cell;
  const float rx=float(xy.x-d.ncenter[0]), ry=float(xy.y-d.ncenter[1]), rz=float(z-d.ncenter[2]);
  vr.x=d.nvel[0]+(d.nomega[1]*rz-d.nomega[2]*ry);
  vr.y=d.nvel[1]+(d.nomega[2]*rx-d.nomega[0]*rz);
  vr.z=d.nvel[2]+(d.nomega[0]*ry-d.nomega[1]*rx);
  velrho[p]=vr;
}

__global__ void k_ft_commit(FtDev* fts, uint32_t ftcount, const StepState* __restrict__ st, double ftpause){
  const uint32_t cf=blockIdx.x*blockDim.x+threadIdx.x;
  if(cf>=ftcount || st->timestep<ftpause)return;
  FtDev& d=fts[cf];
  for(int k=0;k<3;k++){ d.f.center[k]=d.ncenter[k]; d.f.fvel[k]=d.nvel[k]; d.f.fomega[k]=d.nomega[k]; }
}

// phase 1: sums of this context's particles into ft_allsums[6*ftcount]; phase 2: rigid-body algebra + particle update
static int ft_sums(b2s_ctx* ctx){
  cudaStream_t s=ctx->stream;
  FtDev* fts=(FtDev*)ctx->ft_dev;
  for(uint32_t cf=0;cf<ctx->ftcount;cf++){
    const uint32_t nb=b2s_blocks(ctx->fts[cf].count,FT_BS);
    k_ft_sum_blocks<<<nb,FT_BS,0,s>>>(fts,cf,ctx->div,ctx->ridp,ctx->prm.casenfixed,ctx->cur.posxy,ctx->cur.posz,ctx->cur.code,ctx->acear,ctx->ft_sums_dev); B2S_LAUNCHED(ctx);
    k_ft_reduce<<<1,32,0,s>>>(cf,nb,ctx->ft_sums_dev,ctx->ft_allsums_dev); B2S_LAUNCHED(ctx);
  }
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
static int ft_apply(b2s_ctx* ctx, bool predictor){
  cudaStream_t s=ctx->stream;
  FtDev* fts=(FtDev*)ctx->ft_dev;
  const double ftpause=ctx->prm.ftpause;
  for(uint32_t cf=0;cf<ctx->ftcount;cf++){
    const uint32_t nb=b2s_blocks(ctx->fts[cf].count,FT_BS);
    k_ft_compute<<<1,32,0,s>>>(fts,cf,ctx->ft_allsums_dev,ctx->step,predictor? 1: 0,ftpause,ctx->cte); B2S_LAUNCHED(ctx);
    k_ft_update<<<nb,FT_BS,0,s>>>(fts,cf,ctx->div,ctx->step,predictor? 1: 0,ftpause,ctx->ridp,ctx->prm.casenfixed,ctx->cur.posxy,ctx->cur.posz,
      ctx->cur.velrho,ctx->cur.dcell,ctx->cur.code,ctx->cte); B2S_LAUNCHED(ctx);
  }
  if(!predictor){ k_ft_commit<<<b2s_blocks(ctx->ftcount,64),64,0,s>>>(fts,ctx->ftcount,ctx->step,ftpause); B2S_LAUNCHED(ctx); }
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
int b2s_floating_impl(b2s_ctx* ctx, double, bool predictor, bool){
  if(!ctx->ftcount)return B2S_OK;
  int r=ft_sums(ctx); if(r)return r;
  return ft_apply(ctx,predictor);
}
// Slab mode (include/b2s.h): the sums of every rank are all-reduced (SUM) between the two calls, every rank then does the
// same rigid-body algebra on the same numbers and moves the floating particles it owns.
extern "C" int b2s_slab_ft_sums(b2s_ctx* ctx, void** dev_f32, uint32_t* count){
  if(!ctx || !dev_f32 || !count)return B2S_ERR_ARG;
  *dev_f32=NULL; *count=0;
  if(!ctx->ftcount)return B2S_OK;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  int r=ft_sums(ctx); if(r)return r;
  *dev_f32=(void*)ctx->ft_allsums_dev; *count=6u*ctx->ftcount;
  return B2S_OK;
}
extern "C" int b2s_slab_ft_apply(b2s_ctx* ctx, int predictor){
  if(!ctx)return B2S_ERR_ARG;
  if(!ctx->ftcount)return B2S_OK;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  return ft_apply(ctx,predictor!=0);
}

extern "C" int b2s_set_floatings(b2s_ctx* ctx, uint32_t ftcount, const b2s_floating* fin){
  if(!ctx || (ftcount && !fin))return B2S_ERR_ARG;
  if(ftcount>2047)return b2s_fail(ctx,B2S_ERR_ARG,"too many floating bodies");
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  cudaFree(ctx->ft_dev); cudaFree(ctx->ft_massp_dev); cudaFree(ctx->ft_sums_dev); cudaFree(ctx->ft_allsums_dev);
  ctx->ft_dev=NULL; ctx->ft_massp_dev=NULL; ctx->ft_sums_dev=NULL; ctx->ft_allsums_dev=NULL;
  ctx->ftcount=ftcount; ctx->fts.assign(fin,fin+ftcount);
  if(!ftcount)return B2S_OK;
  std::vector<FtDev> h(ftcount); std::vector<float> massp(ftcount);
  uint32_t maxcount=1;
  for(uint32_t k=0;k<ftcount;k++){
    memset(&h[k],0,sizeof(FtDev)); h[k].f=fin[k]; massp[k]=fin[k].massp;
    for(int a=0;a<3;a++){ h[k].ncenter[a]=fin[k].center[a]; h[k].nvel[a]=fin[k].fvel[a]; h[k].nomega[a]=fin[k].fomega[a]; }
    maxcount=std::max(maxcount,fin[k].count);
    if(fin[k].begin<ctx->prm.casenpb || fin[k].begin+fin[k].count>ctx->prm.casenbound)return b2s_fail(ctx,B2S_ERR_ARG,"floating block %u is outside the floating id range",k);
  }
  B2S_CUDA(ctx,cudaMalloc(&ctx->ft_dev,sizeof(FtDev)*ftcount));
  B2S_CUDA(ctx,cudaMalloc((void**)&ctx->ft_massp_dev,sizeof(float)*ftcount));
  B2S_CUDA(ctx,cudaMalloc((void**)&ctx->ft_sums_dev,sizeof(float)*6*(size_t(b2s_blocks(maxcount,FT_BS))+1)));
  B2S_CUDA(ctx,cudaMalloc((void**)&ctx->ft_allsums_dev,sizeof(float)*6*size_t(ftcount)));
  B2S_CUDA(ctx,cudaMemcpy(ctx->ft_dev,h.data(),sizeof(FtDev)*ftcount,cudaMemcpyHostToDevice));
  B2S_CUDA(ctx,cudaMemcpy(ctx->ft_massp_dev,massp.data(),sizeof(float)*ftcount,cudaMemcpyHostToDevice));
  return B2S_OK;
}

extern "C" int b2s_get_floatings(b2s_ctx* ctx, uint32_t ftcount, b2s_floating* out){
  if(!ctx || !out || ftcount>ctx->ftcount)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  std::vector<FtDev> h(ctx->ftcount);
  if(ctx->ftcount)B2S_CUDA(ctx,cudaMemcpy(h.data(),ctx->ft_dev,sizeof(FtDev)*ctx->ftcount,cudaMemcpyDeviceToHost));
  for(uint32_t k=0;k<ftcount;k++){ out[k]=h[k].f; for(int a=0;a<3;a++){ out[k].facelin[a]=h[k].f.facelin[a]; } }
  return B2S_OK;
}
