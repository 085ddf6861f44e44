// misc.cu -- boundary motion, pair counting (PIPS) and small utilities.
//
//   cusph::MoveLinBound (JSphGpu_ker.cu:1671-1700) driven by JSphGpu::RunMotion (JSphGpu.cpp:1173-1225);
//   the motion law is JMotionMovRectSinu (JMotionMov.h / JMotionObj.cpp): pos(t)=ampl*sin(2*pi*freq*t+phase).
//   JDsPips pair counts (JDsPips_ker.cu, JDsPips.cpp:116-142).
#include "b2s_internal.h"
#include <math.h>

// ---- moving boundary ----------------------------------------------------------------------------
struct MotionDev { uint32_t idbegin, count; double freq[3], ampl[3], phase[3], tstart, tend; };

// Displacement over [t,t+dt] and velocity = displacement/dt, like JDsMotion::ProcesTime for MOTT_Linear
// (JDsMotion.cpp: mvpos=pos(t+dt)-pos(t), mvvel=mvpos/dt).  One thread per particle of the block.
__global__ void k_move_rectsinu(MotionDev m, const StepState* __restrict__ st, const DivState* __restrict__ ds, const uint32_t* __restrict__ ridp,
  uint32_t casenfixed, double2* posxy, double* posz, float4* velrho, uint32_t* dcell, typecode* code, const __grid_constant__ DevCte c)
{
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=m.count)return;
  const double t0=st->timestep, dt=st->dt;
  if(t0+dt<=m.tstart || t0>=m.tend)return;
  const double ta=fmax(t0,m.tstart)-m.tstart, tb=fmin(t0+dt,m.tend)-m.tstart;
  const double twopi=6.283185307179586476925;
  double mv[3];
  for(int k=0;k<3;k++)mv[k]=m.ampl[k]*(sin(twopi*m.freq[k]*tb+m.phase[k])-sin(twopi*m.freq[k]*ta+m.phase[k]));
  const float vx=float(mv[0]/dt), vy=float(mv[1]/dt), vz=float(mv[2]/dt);
  const uint32_t p=ridp[m.idbegin-casenfixed+i];
  if(p>=ds->np)return;
  // KerMoveLinBound: pos+=mvpos (double), cell recomputed, vel=mvvel, rho kept.
  double2 xy=posxy[p]; double z=posz[p];
  xy.x+=mv[0]; xy.y+=mv[1]; z+=mv[2];
  uint32_t cell;
  const bool out=b2s_pos_to_cell(c,xy.x,xy.y,z,cell);
  posxy[p]=xy; posz[p]=z;
  if(out)code[p]=typecode(B2S_CODE_SETNORMAL(code[p])|B2S_CODE_OUTPOS);
  dcell[p]=cell;
  float4 v=velrho[p]; v.x=vx; v.y=vy; v.z=vz; velrho[p]=v;
}

int b2s_motion_impl(b2s_ctx* ctx, double, double){
  for(size_t k=0;k<ctx->motions.size();k++){
    const Motion& h=ctx->motions[k];
    MotionDev m; m.idbegin=h.idbegin; m.count=h.count; m.tstart=h.tstart; m.tend=h.tend;
    for(int a=0;a<3;a++){ m.freq[a]=h.freq[a]; m.ampl[a]=h.ampl[a]; m.phase[a]=h.phase[a]; }
    k_move_rectsinu<<<b2s_blocks(m.count,256),256,0,ctx->stream>>>(m,ctx->step,ctx->div,ctx->ridp,ctx->prm.casenfixed,
      ctx->cur.posxy,ctx->cur.posz,ctx->cur.velrho,ctx->cur.dcell,ctx->cur.code,ctx->cte); B2S_LAUNCHED(ctx);
  }
  ctx->bound_changed=true;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}

extern "C" int b2s_add_motion_rectsinu(b2s_ctx* ctx, uint32_t idbegin, uint32_t count, const double freq[3],
                            const double ampl[3], const double phase[3], double tstart, double tend)
{
  if(!ctx || !freq || !ampl || !phase)return B2S_ERR_ARG;
  if(idbegin<ctx->prm.casenfixed || idbegin+count>ctx->prm.casenfixed+ctx->prm.casenmoving)
    return b2s_fail(ctx,B2S_ERR_ARG,"motion block [%u,%u) is not inside the moving id range",idbegin,idbegin+count);
  Motion m; m.idbegin=idbegin; m.count=count; m.tstart=tstart; m.tend=tend;
  for(int k=0;k<3;k++){ m.freq[k]=freq[k]; m.ampl[k]=ampl[k]; m.phase[k]=phase[k]; }
  ctx->motions.push_back(m);
  return B2S_OK;
}

__global__ void k_move_linear(uint32_t idbegin,uint32_t count,double mx,double my,double mz,float vx,float vy,float vz,const DivState* __restrict__ ds,
  const uint32_t* __restrict__ ridp,uint32_t casenfixed,double2* posxy,double* posz,float4* velrho,uint32_t* dcell,typecode* code,const __grid_constant__ DevCte c)
{
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=count)return;
  const uint32_t p=ridp[idbegin-casenfixed+i];
  if(p>=ds->np)return;
  double2 xy=posxy[p]; double z=posz[p];
  xy.x+=mx; xy.y+=my; z+=mz;
  uint32_t cell;
  const bool out=b2s_pos_to_cell(c,xy.x,xy.y,z,cell);
  posxy[p]=xy; posz[p]=z;
  if(out)code[p]=typecode(B2S_CODE_SETNORMAL(code[p])|B2S_CODE_OUTPOS);
  dcell[p]=cell;
  float4 v=velrho[p]; v.x=vx; v.y=vy; v.z=vz; velrho[p]=v;
}

extern "C" int b2s_move_linear(b2s_ctx* ctx, uint32_t idbegin, uint32_t count, const double mvpos[3], const float mvvel[3]){
  if(!ctx || !mvpos || !mvvel)return B2S_ERR_ARG;
  if(idbegin<ctx->prm.casenfixed || idbegin+count>ctx->prm.casenfixed+ctx->prm.casenmoving)
    return b2s_fail(ctx,B2S_ERR_ARG,"block [%u,%u) is not inside the moving id range",idbegin,idbegin+count);
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  k_move_linear<<<b2s_blocks(count,256),256,0,ctx->stream>>>(idbegin,count,mvpos[0],mvpos[1],mvpos[2],mvvel[0],mvvel[1],mvvel[2],ctx->div,ctx->ridp,
    ctx->prm.casenfixed,ctx->cur.posxy,ctx->cur.posz,ctx->cur.velrho,ctx->cur.dcell,ctx->cur.code,ctx->cte); B2S_LAUNCHED(ctx);
  ctx->bound_changed=true;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}

// ---- pair counting (flop model) -----------------------------------------------------------------
__global__ void __launch_bounds__(128) k_count_pairs(const DivState* __restrict__ ds, StepState* st, const uint32_t* __restrict__ begt,
  const uint32_t* __restrict__ dcell, const float4* __restrict__ posp, const __grid_constant__ DevCte c)
{
  const uint32_t p1=ds->npb+blockIdx.x*blockDim.x+threadIdx.x;
  unsigned long long chk=0, real=0;
  if(p1<ds->np){
    const float4 pp1=posp[p1];
    const uint32_t dc=dcell[p1];
    const int cx=int(B2S_DCEL_CELLX(c.domcellcode,dc))-int(ds->cellmin[0]);
    const int cy=int(B2S_DCEL_CELLY(c.domcellcode,dc))-int(ds->cellmin[1]);
    const int cz=int(B2S_DCEL_CELLZ(c.domcellcode,dc))-int(ds->cellmin[2]);
    const int d=c.scelldiv, ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
    const int xini=cx-min(cx,d), xfin=cx+min(ncx-cx-1,d)+1, yini=cy-min(cy,d), yfin=cy+min(ncy-cy-1,d)+1, zini=cz-min(cz,d), zfin=cz+min(ncz-cz-1,d)+1;
    for(int pass=0;pass<2;pass++){
      const uint32_t* beg=begt+(pass==0? ds->box_fluid: 0u);
      for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++)for(int x=xini;x<xfin;x++){
        const uint32_t pini=beg[y*ncx+z*nsheet+x], pfin=beg[y*ncx+z*nsheet+x+1];
        const float ox=b2s_offx(c,cx+int(ds->cellmin[0]),x+int(ds->cellmin[0])), oy=c.scell*float(cy-y), oz=c.scell*float(cz-z);
        for(uint32_t p2=pini;p2<pfin;p2++){
          const float4 pp2=posp[p2];
          const float drx=__fsub_rn(__fadd_rn(pp1.x,ox),pp2.x), dry=__fsub_rn(__fadd_rn(pp1.y,oy),pp2.y), drz=__fsub_rn(__fadd_rn(pp1.z,oz),pp2.z);
          const float rr2=drx*drx+dry*dry+drz*drz;
          chk++;
          if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO)real++;
        }
      }
    }
  }
  for(int o=16;o;o>>=1){ chk+=__shfl_xor_sync(0xffffffffu,chk,o); real+=__shfl_xor_sync(0xffffffffu,real,o); }
  if((threadIdx.x&31)==0){ atomicAdd(&st->pairs_checked,chk); atomicAdd(&st->pairs_real,real); }
}
__global__ void k_reset_pairs(StepState* st){ st->pairs_checked=0; st->pairs_real=0; }

extern "C" int b2s_count_pairs(b2s_ctx* ctx, uint64_t* checked, uint64_t* real){
  if(!ctx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  const uint32_t npf=(ctx->np_host>ctx->npb_host? ctx->np_host-ctx->npb_host: 0u);
  k_reset_pairs<<<1,1,0,ctx->stream>>>(ctx->step); B2S_LAUNCHED(ctx);
  if(npf){ k_count_pairs<<<b2s_blocks(npf,128),128,0,ctx->stream>>>(ctx->div,ctx->step,(const uint32_t*)ctx->begincell,ctx->cur.dcell,ctx->posp,ctx->cte); B2S_LAUNCHED(ctx); }
  int r=b2s_sync_info(ctx); if(r)return r;
  if(checked)*checked=ctx->step_h->pairs_checked;
  if(real)*real=ctx->step_h->pairs_real;
  return B2S_OK;
}
