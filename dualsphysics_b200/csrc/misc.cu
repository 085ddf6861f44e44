// misc.cu -- boundary motion, pair counting (PIPS) and small utilities.
//
//   cusph::MoveLinBound (JSphGpu_ker.cu:1671-1700) driven by JSphGpu::RunMotion (JSphGpu.cpp:1173-1225);
//   the motion law is JMotionMovRectSinu (JMotionMov.h / JMotionObj.cpp): pos(t)=ampl*sin(2*pi*freq*t+phase).
//   JDsPips pair counts (JDsPips_ker.cu, JDsPips.cpp:116-142).
#include "b2s_internal.h"
#include <math.h>

// ---- moving boundary ----------------------------------------------------------------------------
// Displacement of a block over the step [t,t+dt] = the movements of the block that are active in that interval, composed in
// list order like JMotionObj::ProcesTime does with its JMotionPos (JMotionObj.cpp:470-600, JMotionPos.cpp:45-58): translations
// add up while no rotation is involved; once a rotation appears the result is an affine map A <- A*R (the newest movement acts
// on the point first).  Every movement is evaluated in closed form from the time since its start (the reference integrates
// phase / velocity incrementally, step by step: the same function of time).  Purely linear result -> KerMoveLinBound
// (pos+=mov, vel=mov/dt); with a rotation -> KerMoveMatBound (pos=A*pos, vel=(A*pos-pos)/dt, normals carried along)
// (JSphCpu.cpp:2105-2173, JSphGpu_ker.cu:1671-1759).  One thread per particle of the block.
struct Affine { double l[3][3], t[3]; bool rot; };
__device__ __forceinline__ void aff_move(Affine& a,const double d[3]){   // A <- A*Trans(d)
  for(int i=0;i<3;i++)a.t[i]+=a.l[i][0]*d[0]+a.l[i][1]*d[1]+a.l[i][2]*d[2];
}
// rotation by `angdeg` about the axis p1->p2 with the reference's sense (JMatrix4::MatrixRot(ang,p1,p2), JMatrix4.h:465-499:
// seen along the axis from p1 to p2 the body turns by -ang in the right-handed sense), as a Rodrigues map about p1
__device__ __forceinline__ void aff_rotate(Affine& a,double angdeg,const double p1[3],const double p2[3]){
  const double th=-angdeg*0.017453292519943295769;
  double u[3]={p2[0]-p1[0],p2[1]-p1[1],p2[2]-p1[2]};
  const double L=sqrt(u[0]*u[0]+u[1]*u[1]+u[2]*u[2]);
  u[0]/=L; u[1]/=L; u[2]/=L;
  const double c=cos(th), s=sin(th), k=1.-c;
  const double R[3][3]={{c+k*u[0]*u[0],     k*u[0]*u[1]-s*u[2], k*u[0]*u[2]+s*u[1]},
                        {k*u[1]*u[0]+s*u[2], c+k*u[1]*u[1],     k*u[1]*u[2]-s*u[0]},
                        {k*u[2]*u[0]-s*u[1], k*u[2]*u[1]+s*u[0], c+k*u[2]*u[2]}};
  // M = Trans(p1)*R*Trans(-p1): linear part R, translation p1-R*p1;  A <- A*M
  double mt[3]; for(int i=0;i<3;i++)mt[i]=p1[i]-(R[i][0]*p1[0]+R[i][1]*p1[1]+R[i][2]*p1[2]);
  double nl[3][3];
  for(int i=0;i<3;i++)for(int j=0;j<3;j++)nl[i][j]=a.l[i][0]*R[0][j]+a.l[i][1]*R[1][j]+a.l[i][2]*R[2][j];
  for(int i=0;i<3;i++)a.t[i]+=a.l[i][0]*mt[0]+a.l[i][1]*mt[1]+a.l[i][2]*mt[2];
  for(int i=0;i<3;i++)for(int j=0;j<3;j++)a.l[i][j]=nl[i][j];
  a.rot=true;
}

__global__ void k_move_block(const Motion* __restrict__ movs, uint32_t nmov, const StepState* __restrict__ st, const DivState* __restrict__ ds,
  const uint32_t* __restrict__ ridp, uint32_t casenfixed, double2* posxy, double* posz, float4* velrho, uint32_t* dcell, typecode* code,
  float4* boundnor, float4* motionace, const __grid_constant__ DevCte c)
{
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  const Motion& m0=movs[0];
  if(i>=m0.count)return;
  const double t0=st->timestep, dt=st->dt;
  const double twopi=6.283185307179586476925;
  Affine a; for(int r=0;r<3;r++){ for(int q=0;q<3;q++)a.l[r][q]=(r==q? 1.: 0.); a.t[r]=0; } a.rot=false;
  bool modif=false;
  for(uint32_t k=0;k<nmov;k++){
    const Motion& m=movs[k];
    if(t0+dt<=m.tstart || t0>=m.tend)continue;
    const double ta=fmax(t0,m.tstart)-m.tstart, tb=fmin(t0+dt,m.tend)-m.tstart;   // movement-relative times of the overlap
    if(!(tb>ta))continue;
    double d[3]={0,0,0};
    switch(m.type){
      case B2S_MOV_RECTSINU: for(int q=0;q<3;q++)d[q]=m.v2[q]*(sin(twopi*m.v1[q]*tb+m.v3[q])-sin(twopi*m.v1[q]*ta+m.v3[q])); aff_move(a,d); modif=true; break;
      case B2S_MOV_RECT:     for(int q=0;q<3;q++)d[q]=m.v1[q]*(tb-ta); aff_move(a,d); modif=true; break;
      case B2S_MOV_RECTACE:  for(int q=0;q<3;q++)d[q]=m.v1[q]*(tb-ta)+0.5*m.v2[q]*(tb*tb-ta*ta); aff_move(a,d); modif=true; break;
      case B2S_MOV_ROT:      aff_rotate(a,m.v1[0]*(tb-ta),m.p1,m.p2); modif=true; break;
      case B2S_MOV_ROTACE:   aff_rotate(a,m.v1[0]*(tb-ta)+0.5*m.v2[0]*(tb*tb-ta*ta),m.p1,m.p2); modif=true; break;
      case B2S_MOV_ROTSINU:  aff_rotate(a,m.v2[0]*(sin(twopi*m.v1[0]*tb+m.v3[0])-sin(twopi*m.v1[0]*ta+m.v3[0])),m.p1,m.p2); modif=true; break;
      default: break;   // B2S_MOV_WAIT
    }
  }
  const uint32_t p=ridp[m0.idbegin-casenfixed+i];
  if(p>=ds->np)return;
  if(!modif){   // no active movement: position and velocity stay (JSphCpu::RunMotion only touches MOTT_Linear / MOTT_Matrix objects)
    if(motionace)motionace[p]=make_float4(0.f,0.f,0.f,0.f);   // CopyMotionVelAce still runs: the velocity did not change
    return;
  }
  double2 xy=posxy[p]; double z=posz[p];
  double mv[3];
  if(!a.rot){ mv[0]=a.t[0]; mv[1]=a.t[1]; mv[2]=a.t[2]; }
  else{
    const double nx=a.l[0][0]*xy.x+a.l[0][1]*xy.y+a.l[0][2]*z+a.t[0], ny=a.l[1][0]*xy.x+a.l[1][1]*xy.y+a.l[1][2]*z+a.t[1], nz=a.l[2][0]*xy.x+a.l[2][1]*xy.y+a.l[2][2]*z+a.t[2];
    mv[0]=nx-xy.x; mv[1]=ny-xy.y; mv[2]=nz-z;
    if(boundnor){   // the normal is carried as the image of the ghost node (JSphCpu.cpp:2166-2170)
      const float4 nr=boundnor[p];
      const double gx=xy.x+double(nr.x), gy=xy.y+double(nr.y), gz=z+double(nr.z);
      const double hx=a.l[0][0]*gx+a.l[0][1]*gy+a.l[0][2]*gz+a.t[0], hy=a.l[1][0]*gx+a.l[1][1]*gy+a.l[1][2]*gz+a.t[1], hz=a.l[2][0]*gx+a.l[2][1]*gy+a.l[2][2]*gz+a.t[2];
      boundnor[p]=make_float4(float(hx-nx),float(hy-ny),float(hz-nz),nr.w);
    }
  }
  const float vx=float(mv[0]/dt), vy=float(mv[1]/dt), vz=float(mv[2]/dt);
  // KerMoveLinBound / KerMoveMatBound: pos+=mov (double), cell recomputed, vel=mov/dt, rho kept.
  xy.x+=mv[0]; xy.y+=mv[1]; z+=mv[2];
  uint32_t cell;
  const bool out=b2s_pos_to_cell(c,xy.x,xy.y,z,cell);
  posxy[p]=xy; posz[p]=z;
  if(out)code[p]=typecode(B2S_CODE_SETNORMAL(code[p])|B2S_CODE_OUTPOS);
  dcell[p]=cell;
  float4 v=velrho[p];
  // CopyMotionVelAce (JSphCpu_mdbc.cpp:472-490): acceleration of the imposed motion = change of the motion velocity over the step
  if(motionace)motionace[p]=make_float4(float((double(vx)-double(v.x))/dt),float((double(vy)-double(v.y))/dt),float((double(vz)-double(v.z))/dt),0.f);
  v.x=vx; v.y=vy; v.z=vz; velrho[p]=v;
}

int b2s_motion_impl(b2s_ctx* ctx, double, double){
  const size_t n=ctx->motions.size();
  if(!n)return B2S_OK;
  if(ctx->motion_dirty){
    cudaFree(ctx->motion_dev); ctx->motion_dev=NULL;
    B2S_CUDA(ctx,cudaMalloc(&ctx->motion_dev,n*sizeof(Motion)));
    B2S_CUDA(ctx,cudaMemcpyAsync(ctx->motion_dev,ctx->motions.data(),n*sizeof(Motion),cudaMemcpyHostToDevice,ctx->stream));
    ctx->motion_dirty=false;
  }
  // one launch per block of particles: its movements are contiguous in the list (b2s_add_motion keeps them together)
  for(size_t k=0;k<n;){
    size_t e=k+1;
    while(e<n && ctx->motions[e].idbegin==ctx->motions[k].idbegin)e++;
    const Motion& h=ctx->motions[k];
    k_move_block<<<b2s_blocks(h.count,256),256,0,ctx->stream>>>((const Motion*)ctx->motion_dev+k,uint32_t(e-k),ctx->step,ctx->div,ctx->ridp,ctx->prm.casenfixed,
      ctx->cur.posxy,ctx->cur.posz,ctx->cur.velrho,ctx->cur.dcell,ctx->cur.code,ctx->boundnor,ctx->motionace,ctx->cte); B2S_LAUNCHED(ctx);
    k=e;
  }
  ctx->bound_changed=true;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}

extern "C" int b2s_add_motion(b2s_ctx* ctx, const b2s_motion* mi){
  if(!ctx || !mi)return B2S_ERR_ARG;
  if(mi->type<B2S_MOV_RECTSINU || mi->type>B2S_MOV_WAIT)return b2s_fail(ctx,B2S_ERR_ARG,"motion type %d is unknown",mi->type);
  if(mi->idbegin<ctx->prm.casenfixed || mi->idbegin+mi->count>ctx->prm.casenfixed+ctx->prm.casenmoving)
    return b2s_fail(ctx,B2S_ERR_ARG,"motion block [%u,%u) is not inside the moving id range",mi->idbegin,mi->idbegin+mi->count);
  const bool isrot=(mi->type==B2S_MOV_ROT || mi->type==B2S_MOV_ROTACE || mi->type==B2S_MOV_ROTSINU);
  if(isrot && mi->axisp1[0]==mi->axisp2[0] && mi->axisp1[1]==mi->axisp2[1] && mi->axisp1[2]==mi->axisp2[2])
    return b2s_fail(ctx,B2S_ERR_ARG,"rotation axis: axisp1 and axisp2 coincide");
  Motion m; memset(&m,0,sizeof(m));
  m.idbegin=mi->idbegin; m.count=mi->count; m.type=mi->type; m.tstart=mi->tstart; m.tend=mi->tend;
  for(int k=0;k<3;k++){ m.v1[k]=mi->v1[k]; m.v2[k]=mi->v2[k]; m.v3[k]=mi->v3[k]; m.p1[k]=mi->axisp1[k]; m.p2[k]=mi->axisp2[k]; }
  // keep the movements of one block together, in the order they were added
  size_t pos=ctx->motions.size();
  for(size_t k=0;k<ctx->motions.size();k++)if(ctx->motions[k].idbegin==m.idbegin)pos=k+1;
  ctx->motions.insert(ctx->motions.begin()+pos,m);
  ctx->motion_dirty=true;
  return B2S_OK;
}

extern "C" int b2s_add_motion_rectsinu(b2s_ctx* ctx, uint32_t idbegin, uint32_t count, const double freq[3],
                            const double ampl[3], const double phase[3], double tstart, double tend)
{
  if(!ctx || !freq || !ampl || !phase)return B2S_ERR_ARG;
  b2s_motion m; memset(&m,0,sizeof(m));
  m.idbegin=idbegin; m.count=count; m.type=B2S_MOV_RECTSINU; m.tstart=tstart; m.tend=tend;
  for(int k=0;k<3;k++){ m.v1[k]=freq[k]; m.v2[k]=ampl[k]; m.v3[k]=phase[k]; }
  return b2s_add_motion(ctx,&m);
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
  const uint32_t npf=ctx->npf_host;
  k_reset_pairs<<<1,1,0,ctx->stream>>>(ctx->step); B2S_LAUNCHED(ctx);
  if(npf){ k_count_pairs<<<b2s_blocks(npf,128),128,0,ctx->stream>>>(ctx->div,ctx->step,(const uint32_t*)ctx->begincell,ctx->cur.dcell,ctx->posp,ctx->cte); B2S_LAUNCHED(ctx); }
  int r=b2s_sync_info(ctx); if(r)return r;
  if(checked)*checked=ctx->step_h->pairs_checked;
  if(real)*real=ctx->step_h->pairs_real;
  return B2S_OK;
}
