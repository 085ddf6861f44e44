// shifting.cu -- standard particle shifting (Lind et al. as coded by the reference), whole-domain form.
//
// Replaces (under /root/reference/src/source):
//   the shiftposfs accumulators of the pair loop     JSphCpu.cpp:767-779,848,942-950,1101 (GPU: JSphGpu_ker.cu:676-684)
//   JSphShifting::InitCpu (no zones: memset 0)        JSphShifting.cpp:346-350
//   JSphShifting::RunCpu / JSphCpu::RunShifting       JSphShifting.cpp:390-423, JSphCpu.cpp:2084-2088
// callers JSphCpuSingle::ComputeStep_Ver/_Sym (JSphCpuSingle.cpp:881,909,924).
//
// One kernel: every owned fluid particle gathers sum(m/rho * gradW) (xyz) and -sum(m/rho * r.gradW) (w: the
// free-surface divergence) over its fluid neighbours and then over its boundary neighbours -- one running sum in
// the reference's order -- and turns it straight into the displacement of this sub-step, so the float4 sum never
// goes through HBM.  It is kept apart from the three interaction kernel variants so that those stay bit-identical
// with and without shifting; the price is a second neighbour walk (only in shifting runs).
// The Symplectic predictor ignores the displacement (`shift=false`, JSphCpu.cpp:1689), so it is only evaluated
// before the corrector and before the Verlet update; the reference's store rule `if(shift||...)` of the interaction
// itself (JSphCpu.cpp:1084) is honoured by finalize_store (forces_common.cuh) in both sub-steps.
#include "forces_common.cuh"

struct ShiftArgs {
  const DivState* ds; const StepState* st;
  const uint32_t* beg; const uint32_t* dcell; const typecode* code;
  const float4* posp; const float4* velrho;
  const float* ftomassp; const uint8_t* boundmode;
  float4* shiftpos;
};

template<int TKER,bool FLOATING,bool BOUNDP2>
__device__ __forceinline__ void shift_rows(const ShiftArgs& g,const DevCte& c,const CellRange& r,const float4 pp1,float xin,float4& s){
  const DivState* ds=g.ds;
  const int ncx=int(ds->ncx), nsheet=int(ds->nsheet), cmx=int(ds->cellmin[0]);
  const uint32_t* beg=(BOUNDP2? g.beg: g.beg+ds->box_fluid);
  for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
    const int row=y*ncx+z*nsheet;
    const float offy=c.scell*float(r.cy-y), offz=c.scell*float(r.cz-z);
    int xa,xb;
    if(!row_trim(c,r,xin,FADD(pp1.y,offy),FADD(pp1.z,offz),xa,xb))continue;
    for(int x=xa;x<xb;x++){
      const uint32_t pini=beg[row+x], pfin=beg[row+x+1];
      const float offx=b2s_offx(c,r.cx+cmx,x+cmx);
      for(uint32_t p2=pini;p2<pfin;p2++){
        const float4 pp2=g.posp[p2];
        const float drx=b2s_dr(pp1.x,pp2.x,offx), dry=b2s_dr(pp1.y,pp2.y,offy), drz=b2s_dr(pp1.z,pp2.z,offz);
        const float rr2=b2s_rr2(drx,dry,drz);
        if(rr2<=c.kernelsize2 && rr2>=B2S_ALMOSTZERO){
          const float fac=kernel_fac<TKER>(rr2,c);
          const float frx=FMUL(fac,drx), fry=FMUL(fac,dry), frz=FMUL(fac,drz);
          float massp2=(BOUNDP2? c.massb: c.massf);
          typecode cod2=0;
          if(FLOATING || (BOUNDP2 && c.shiftmode==B2S_SHIFT_NOFIXED))cod2=g.code[p2];
          bool ftp2=false;
          if(FLOATING){
            ftp2=B2S_CODE_ISFLOATING(cod2);
            if(ftp2){ massp2=g.ftomassp[B2S_CODE_TYPEVALUE(cod2)]; if(c.shiftmode==B2S_SHIFT_NOBOUND)s.x=FLT_MAX; }   // JSphCpu.cpp:848
          }
          if(BOUNDP2 && c.mdbc2 && !ftp2 && g.boundmode[p2]==B2S_BMODE_MDBC2OFF)massp2=0.f;                        // JSphCpu.cpp:851-853
          if(s.x!=FLT_MAX){                                                                                          // JSphCpu.cpp:943-950
            const float massrho=FMUL(massp2,b2s_rcp(g.velrho[p2].w));
            const bool noshift=(BOUNDP2 && (c.shiftmode==B2S_SHIFT_NOBOUND || (c.shiftmode==B2S_SHIFT_NOFIXED && B2S_CODE_TYPE(cod2)==B2S_CODE_TYPE_FIXED)));
            s.x=(noshift? FLT_MAX: FFMA(massrho,frx,s.x));
            s.y=FFMA(massrho,fry,s.y);
            s.z=FFMA(massrho,frz,s.z);
            s.w=FFMA(-massrho,b2s_dot3(drx,dry,drz,frx,fry,frz),s.w);
          }
        }
      }
    }
  }
}

template<int TKER,bool FLOATING>
__global__ void __launch_bounds__(F_BS) k_shifting(ShiftArgs g, const __grid_constant__ DevCte c)
{
  const DivState* ds=g.ds;
  const uint32_t p1=ds->npb+blockIdx.x*blockDim.x+threadIdx.x;
  if(p1>=ds->np)return;
  const typecode cod1=g.code[p1];
  if(!B2S_CODE_ISNORMAL(cod1)){ g.shiftpos[p1]=make_float4(0.f,0.f,0.f,0.f); return; }   // slab ghosts: neighbours only
  float4 s=make_float4(0.f,0.f,0.f,0.f);
  if(FLOATING && B2S_CODE_ISFLOATING(cod1))s.x=FLT_MAX;                                     // JSphCpu.cpp:778
  else{
    const float4 pp1=g.posp[p1];
    const CellRange r=cell_range(g.dcell[p1],c,ds);
    const float xin=pp1.x-c.scell*float((r.cx+int(ds->cellmin[0]))&((1<<B2S_XBLK)-1));
    shift_rows<TKER,FLOATING,false>(g,c,r,pp1,xin,s);
    shift_rows<TKER,FLOATING,true >(g,c,r,pp1,xin,s);
  }
  //-JSphShifting::RunCpu (JSphShifting.cpp:390-423).
  float4 res=make_float4(0.f,0.f,0.f,0.f);
  if(s.x!=FLT_MAX){
    const double coefumagn=g.st->dt*double(c.shiftcoef)*double(c.kernelh);
    const float4 v=g.velrho[p1];
    const double vx=double(v.x), vy=double(v.y), vz=double(v.z);
    double umagn=coefumagn*sqrt(vx*vx+vy*vy+vz*vz);
    if(c.shifttfs!=0.f){
      const double coeftfs=(c.simulate2d? 2.0: 3.0)-double(c.shifttfs);
      if(s.w<c.shifttfs)umagn=0;
      else umagn*=(double(s.w)-double(c.shifttfs))/coeftfs;
    }
    const float dx=float(double(s.x)*umagn), dy=float(double(s.y)*umagn), dz=float(double(s.z)*umagn);
    const float md=c.shiftmaxdist;
    res.x=(fabsf(dx)<md? dx: (dx>=0? md: -md));
    res.y=(fabsf(dy)<md? dy: (dy>=0? md: -md));
    res.z=(fabsf(dz)<md? dz: (dz>=0? md: -md));
    res.w=s.w;   // RunCpu leaves the divergence in place
  }
  g.shiftpos[p1]=res;
}

// Displacement of the coming position update; StepState.dt must already hold the dt of the sub-step.
int b2s_shifting_impl(b2s_ctx* ctx){
  if(!ctx->prm.shiftmode)return B2S_OK;
  const uint32_t npf=ctx->npf_host;
  if(!npf)return B2S_OK;
  ShiftArgs g;
  g.ds=ctx->div; g.st=ctx->step; g.beg=(const uint32_t*)ctx->begincell; g.dcell=ctx->cur.dcell; g.code=ctx->cur.code;
  g.posp=ctx->posp; g.velrho=ctx->cur.velrho; g.ftomassp=ctx->ft_massp_dev; g.boundmode=ctx->boundmode; g.shiftpos=ctx->shiftpos;
  const bool ft=(ctx->cte.use_floating!=0);
  const unsigned nb=b2s_blocks(npf,F_BS);
  if(ctx->cte.tkernel==B2S_KERNEL_WENDLAND){
    if(ft)k_shifting<B2S_KERNEL_WENDLAND,true><<<nb,F_BS,0,ctx->stream>>>(g,ctx->cte); else k_shifting<B2S_KERNEL_WENDLAND,false><<<nb,F_BS,0,ctx->stream>>>(g,ctx->cte);
  }
  else{
    if(ft)k_shifting<B2S_KERNEL_CUBIC,true><<<nb,F_BS,0,ctx->stream>>>(g,ctx->cte); else k_shifting<B2S_KERNEL_CUBIC,false><<<nb,F_BS,0,ctx->stream>>>(g,ctx->cte);
  }
  B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
