// integrate.cu -- time integration fused with the position/cell update.
//
// Replaces (under /root/reference/src/source):
//   cusphs::ComputeStepVerlet        JSphGpuSimple_ker.cu:132-260   caller JSphGpu::ComputeVerlet        JSphGpu.cpp:939-990
//   cusphs::ComputeStepSymplecticPre JSphGpuSimple_ker.cu:265-390   caller JSphGpu::ComputeSymplecticPre JSphGpu.cpp:996-1048
//   cusphs::ComputeStepSymplecticCor JSphGpuSimple_ker.cu:395-553   caller JSphGpu::ComputeSymplecticCorr JSphGpu.cpp:1054-1097
//   cusph::ComputeStepPos/Pos2 -> KerUpdatePos   JSphGpu_ker.cu:334-392,1476-1570
// The reference runs two kernels (velocity/density + displacement "mov" arrays, then position/cell);
// here one kernel per integrator does both, so the 24 B/particle mov arrays never exist.
// Arithmetic is the CPU statement's (JSphCpu.cpp:1458-1510 UpdatePos, :1523-1602 Verlet, :1687-1850
// SymplecticPre, :1856-2022 SymplecticCorr): every update is float(double(old)+double(rate)*dt); the
// cell hash is unsigned(dx/double(float scell)), never a reciprocal multiply.
// dt is read from the device-resident StepState so a whole step needs no host synchronisation.
#include "b2s_internal.h"

#define I_BS 256

struct IntegArgs {
  const DivState* ds; const StepState* st;
  typecode* code; uint32_t* dcell;
  // current state (read) and destination (write) -- may alias per element
  const double2* posxy_r; const double* posz_r; const float4* velrho_r;
  double2* posxy_w; double* posz_w; float4* velrho_w;
  const float4* velrho_aux;   // verlet: velrho(n-1) (or == velrho on Euler steps); symcorr: predicted velrho (rho*)
  const float4* acear;
  int euler;                  // verlet only
  const float4* shiftpos;     // standard shifting (else NULL): displacement added to the fluid particles' move
  const uint8_t* boundmode;   // mDBC2 only (else NULL): boundary particles with mode>=BMODE_MDBC2 keep their velrho
};

// KerUpdatePos / JSphCpu::UpdatePos (JSphCpu.cpp:1458-1516), periodic wrap included (b2s_pos_to_cell).
__device__ __forceinline__ void update_pos(const DevCte& c, double rx,double ry,double rz, double movx,double movy,double movz,
  bool outrho, typecode rcode, uint32_t p, double2* posxy, double* posz, uint32_t* dcell, typecode* code)
{
  const bool outmov=(fabsf(float(movx))>c.movlimit || fabsf(float(movy))>c.movlimit || fabsf(float(movz))>c.movlimit);
  rx+=movx; ry+=movy; rz+=movz;
  uint32_t cell;
  const bool out=b2s_pos_to_cell(c,rx,ry,rz,cell);
  posxy[p]=make_double2(rx,ry); posz[p]=rz;
  if(outrho || outmov || out){
    if(out)rcode=typecode(B2S_CODE_SETNORMAL(rcode)|B2S_CODE_OUTPOS);
    else if(outrho)rcode=typecode(B2S_CODE_SETNORMAL(rcode)|B2S_CODE_OUTRHO);
    else rcode=typecode(B2S_CODE_SETNORMAL(rcode)|B2S_CODE_OUTMOV);
    code[p]=rcode;
    dcell[p]=B2S_DCEL_MAPOUT;
  }
  else dcell[p]=cell;
}

// ---- Symplectic predictor ---------------------------------------------------------------------
__global__ void __launch_bounds__(I_BS) k_sym_pre(IntegArgs a, const __grid_constant__ DevCte c){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=a.ds->np)return;
  if(B2S_CODE_SPECIAL(a.code[p])==B2S_CODE_PERIODIC){ // slab halo ghost: frozen, dropped by the next cell divide
    a.velrho_w[p]=a.velrho_r[p]; a.posxy_w[p]=a.posxy_r[p]; a.posz_w[p]=a.posz_r[p];
    a.code[p]=typecode(B2S_CODE_SETNORMAL(a.code[p])|B2S_CODE_OUTIGNORE);
    return;
  }
  const double dt05=a.st->dt*.5;
  const float4 pre=a.velrho_r[p];
  const float4 acear=a.acear[p];
  const double2 xy=a.posxy_r[p]; const double z=a.posz_r[p];
  if(p<a.ds->npb){
    float4 vr=pre;
    if(!a.boundmode || a.boundmode[p]<B2S_BMODE_MDBC2){   // DBC / mDBC vel0 (JSphCpu.cpp:1717)
      vr.w=float(double(vr.w)+dt05*double(acear.w));
      vr.w=(vr.w<c.rhopzero? c.rhopzero: vr.w);
    }
    a.velrho_w[p]=vr; a.posxy_w[p]=xy; a.posz_w[p]=z;
    return;
  }
  const typecode rcode=a.code[p];
  const float rhonew=float(double(pre.w)+dt05*double(acear.w));
  if(!c.use_floating || B2S_CODE_ISFLUID(rcode)){
    const double dx=double(pre.x)*dt05, dy=double(pre.y)*dt05, dz=double(pre.z)*dt05;
    const bool outrho=(rhonew<c.rhopoutmin || rhonew>c.rhopoutmax);
    float4 vn;
    vn.x=float(double(pre.x)+(double(acear.x)+double(c.gravity[0]))*dt05);
    vn.y=float(double(pre.y)+(double(acear.y)+double(c.gravity[1]))*dt05);
    vn.z=float(double(pre.z)+(double(acear.z)+double(c.gravity[2]))*dt05);
    vn.w=rhonew;
    a.velrho_w[p]=vn;
    update_pos(c,xy.x,xy.y,z,dx,dy,dz,outrho,rcode,p,a.posxy_w,a.posz_w,a.dcell,a.code);
  }
  else{ // floating: velocity kept, density clamped (JSphCpu.cpp:1786-1790)
    float4 vr=pre; vr.w=(rhonew<c.rhopzero? c.rhopzero: rhonew);
    a.velrho_w[p]=vr; a.posxy_w[p]=xy; a.posz_w[p]=z;
  }
}

// ---- Symplectic corrector ---------------------------------------------------------------------
// reads pre arrays via *_r, predicted density via velrho_aux, writes the current arrays.
__global__ void __launch_bounds__(I_BS) k_sym_corr(IntegArgs a, const __grid_constant__ DevCte c){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=a.ds->np)return;
  if(B2S_CODE_SPECIAL(a.code[p])==B2S_CODE_PERIODIC){ // slab halo ghost: frozen, dropped by the next cell divide
    a.velrho_w[p]=a.velrho_aux[p];
    a.code[p]=typecode(B2S_CODE_SETNORMAL(a.code[p])|B2S_CODE_OUTIGNORE);
    return;
  }
  const double dt=a.st->dt, dt05=dt*.5;
  const float4 pre=a.velrho_r[p];
  const float4 acear=a.acear[p];
  const float rhopred=a.velrho_aux[p].w;
  const double epsilon_rdot=(-double(acear.w)/double(rhopred))*dt;
  const float rhonew=float(double(pre.w)*(2.-epsilon_rdot)/(2.+epsilon_rdot));
  if(p<a.ds->npb){
    if(a.boundmode && a.boundmode[p]>=B2S_BMODE_MDBC2)a.velrho_w[p]=pre;   // mDBC2: velrho of the start of the step (JSphCpu.cpp:1869-1875)
    else a.velrho_w[p]=make_float4(0.f,0.f,0.f,(rhonew<c.rhopzero? c.rhopzero: rhonew));
    return;   // boundary positions: unchanged (pos already equals pospre)
  }
  const typecode rcode=a.code[p];
  if(!c.use_floating || B2S_CODE_ISFLUID(rcode)){
    float4 vn;
    vn.x=float(double(pre.x)+(double(acear.x)+double(c.gravity[0]))*dt);
    vn.y=float(double(pre.y)+(double(acear.y)+double(c.gravity[1]))*dt);
    vn.z=float(double(pre.z)+(double(acear.z)+double(c.gravity[2]))*dt);
    vn.w=rhonew;
    double dx=(double(pre.x)+double(vn.x))*dt05;
    double dy=(double(pre.y)+double(vn.y))*dt05;
    double dz=(double(pre.z)+double(vn.z))*dt05;
    if(a.shiftpos){ const float4 sh=a.shiftpos[p]; dx+=double(sh.x); dy+=double(sh.y); dz+=double(sh.z); }   // JSphCpu.cpp:1936-1940
    const bool outrho=(rhonew<c.rhopoutmin || rhonew>c.rhopoutmax);
    a.velrho_w[p]=vn;
    const double2 xy=a.posxy_r[p]; const double z=a.posz_r[p];
    update_pos(c,xy.x,xy.y,z,dx,dy,dz,outrho,rcode,p,a.posxy_w,a.posz_w,a.dcell,a.code);
  }
  else{
    float4 vr=pre; vr.w=(rhonew<c.rhopzero? c.rhopzero: rhonew);
    a.velrho_w[p]=vr;
    a.posxy_w[p]=a.posxy_r[p]; a.posz_w[p]=a.posz_r[p];
  }
}

// ---- Verlet -------------------------------------------------------------------------------------
// velrho_r = velrho(n), velrho_aux = velrho(n-1) (or velrho(n) on Euler steps); new values -> velrho_w (= the m1 buffer).
__global__ void __launch_bounds__(I_BS) k_verlet(IntegArgs a, const __grid_constant__ DevCte c){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=a.ds->np)return;
  if(B2S_CODE_SPECIAL(a.code[p])==B2S_CODE_PERIODIC){ // slab halo ghost: frozen, dropped by the next cell divide
    a.velrho_w[p]=a.velrho_r[p];
    a.code[p]=typecode(B2S_CODE_SETNORMAL(a.code[p])|B2S_CODE_OUTIGNORE);
    return;
  }
  const double dt=a.st->dt;
  const double dt2=(a.euler? dt: dt+dt);
  const float4 v1=a.velrho_r[p];
  const float4 v2=a.velrho_aux[p];
  const float4 acear=a.acear[p];
  const float rhonew=float(double(v2.w)+dt2*double(acear.w));
  if(p<a.ds->npb){
    if(a.boundmode && a.boundmode[p]>=B2S_BMODE_MDBC2)a.velrho_w[p]=v2;    // mDBC2: ComputeVelrhoBound keeps velrhoold (JSphCpu.cpp:1616)
    else a.velrho_w[p]=make_float4(0.f,0.f,0.f,(rhonew<c.rhopzero? c.rhopzero: rhonew));
    return;
  }
  const typecode rcode=a.code[p];
  if(!c.use_floating || B2S_CODE_ISFLUID(rcode)){
    const double dt205=0.5*dt*dt;
    const double agx=double(acear.x)+double(c.gravity[0]), agy=double(acear.y)+double(c.gravity[1]), agz=double(acear.z)+double(c.gravity[2]);
    double dx=double(v1.x)*dt+agx*dt205, dy=double(v1.y)*dt+agy*dt205, dz=double(v1.z)*dt+agz*dt205;
    if(a.shiftpos){ const float4 sh=a.shiftpos[p]; dx+=double(sh.x); dy+=double(sh.y); dz+=double(sh.z); }   // JSphCpu.cpp:1543-1547
    const bool outrho=(rhonew<c.rhopoutmin || rhonew>c.rhopoutmax);
    float4 vn;
    vn.x=float(double(v2.x)+agx*dt2); vn.y=float(double(v2.y)+agy*dt2); vn.z=float(double(v2.z)+agz*dt2); vn.w=rhonew;
    a.velrho_w[p]=vn;
    const double2 xy=a.posxy_r[p]; const double z=a.posz_r[p];
    update_pos(c,xy.x,xy.y,z,dx,dy,dz,outrho,rcode,p,a.posxy_w,a.posz_w,a.dcell,a.code);
  }
  else{
    float4 vr=v1; vr.w=(rhonew<c.rhopzero? c.rhopzero: rhonew);
    a.velrho_w[p]=vr;
  }
}

__global__ void k_set_dt_host(StepState* st, double dt){ st->dt=dt; }
__global__ void k_set_dt_from(StepState* st, int which){ st->dt=(which==0? st->symplectic_dtpre: st->dt_c); }

int b2s_set_dt_device(b2s_ctx* ctx, int which){ k_set_dt_from<<<1,1,0,ctx->stream>>>(ctx->step,which); B2S_LAUNCHED(ctx); return B2S_OK; }
int b2s_set_dt_host(b2s_ctx* ctx, double dt){ k_set_dt_host<<<1,1,0,ctx->stream>>>(ctx->step,dt); B2S_LAUNCHED(ctx); return B2S_OK; }

// mode: 0 verlet, 1 symplectic predictor, 2 symplectic corrector.  dt must already be in StepState.dt.
int b2s_integrate_impl(b2s_ctx* ctx, int mode, double dt, bool dt_from_device){
  cudaStream_t s=ctx->stream;
  if(!dt_from_device)b2s_set_dt_host(ctx,dt);
  const uint32_t n=ctx->np_host;
  if(!n)return B2S_OK;
  IntegArgs a; memset(&a,0,sizeof(a));
  a.ds=ctx->div; a.st=ctx->step; a.code=ctx->cur.code; a.dcell=ctx->cur.dcell; a.acear=ctx->acear;
  a.boundmode=(ctx->prm.slipmode>=B2S_SLIP_NOSLIP? ctx->boundmode: NULL);
  a.shiftpos=(ctx->prm.shiftmode? ctx->shiftpos: NULL);
  const unsigned g=b2s_blocks(n,I_BS);
  // RunShifting(dt) before the corrector / the Verlet update (JSphCpuSingle.cpp:881,924); the predictor does not use it (JSphCpu.cpp:1689).
  if(a.shiftpos && mode!=1){ const int r=b2s_shifting_impl(ctx); if(r)return r; }
  if(mode==1){
    if(ctx->pre_active)return b2s_fail(ctx,B2S_ERR_STATE,"symplectic predictor called twice");
    // new state -> alt buffers; old state becomes PRE; old PRE buffers become the scratch.
    a.posxy_r=ctx->cur.posxy; a.posz_r=ctx->cur.posz; a.velrho_r=ctx->cur.velrho;
    a.posxy_w=ctx->alt.posxy; a.posz_w=ctx->alt.posz; a.velrho_w=ctx->alt.velrho;
    k_sym_pre<<<g,I_BS,0,s>>>(a,ctx->cte); B2S_LAUNCHED(ctx);
    double2* oxy=ctx->cur.posxy; double* oz=ctx->cur.posz; float4* ov=ctx->cur.velrho;
    ctx->cur.posxy=ctx->alt.posxy; ctx->cur.posz=ctx->alt.posz; ctx->cur.velrho=ctx->alt.velrho;
    ctx->alt.posxy=ctx->pospre_xy; ctx->alt.posz=ctx->pospre_z; ctx->alt.velrho=ctx->velrhopre;
    ctx->pospre_xy=oxy; ctx->pospre_z=oz; ctx->velrhopre=ov;
    ctx->pre_active=true;
  }
  else if(mode==2){
    if(!ctx->pre_active)return b2s_fail(ctx,B2S_ERR_STATE,"symplectic corrector without predictor");
    a.posxy_r=ctx->pospre_xy; a.posz_r=ctx->pospre_z; a.velrho_r=ctx->velrhopre;
    a.velrho_aux=ctx->cur.velrho;
    a.posxy_w=ctx->cur.posxy; a.posz_w=ctx->cur.posz; a.velrho_w=ctx->cur.velrho;
    k_sym_corr<<<g,I_BS,0,s>>>(a,ctx->cte); B2S_LAUNCHED(ctx);
    ctx->pre_active=false;
  }
  else{
    if(!ctx->velrhom1)return b2s_fail(ctx,B2S_ERR_STATE,"verlet arrays not allocated (tstep!=verlet)");
    ctx->verletstep++;   // host-side counter (JSphGpu::ComputeVerlet, JSphGpu.cpp:946-965)
    const bool euler=!(ctx->verletstep<(uint32_t)ctx->prm.verletsteps);
    if(euler)ctx->verletstep=0;
    a.euler=euler? 1: 0;
    a.posxy_r=ctx->cur.posxy; a.posz_r=ctx->cur.posz; a.velrho_r=ctx->cur.velrho;
    a.velrho_aux=(euler? ctx->cur.velrho: ctx->velrhom1);
    a.posxy_w=ctx->cur.posxy; a.posz_w=ctx->cur.posz; a.velrho_w=ctx->velrhom1;
    k_verlet<<<g,I_BS,0,s>>>(a,ctx->cte); B2S_LAUNCHED(ctx);
    std::swap(ctx->cur.velrho,ctx->velrhom1);
  }
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
