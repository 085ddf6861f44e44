// config.cpp -- host-side constant and domain configuration (no CUDA needed).
//
// b2s_config_finalize() restates, from the primary values of b2s_params, what the reference computes in
//   JSph::ConfigConstants1/2        src/source/JSph.cpp:1481-1546
//   fsph::GetKernelWendland_Ctes    src/source/FunSphKernel.h:198-209
//   fsph::GetKernelCubic_Ctes       src/source/FunSphKernel.h:45-85
//   JSph::LoadCaseParticles (map)   src/source/JSph.cpp:2151-2198
//   JSph::ConfigCellDivision        src/source/JSph.cpp:1877-1895
//   JSph::SelecDomain               src/source/JSph.cpp:1901-1946
//   JDsDcell::CalcCellCode          src/source/JDsDcell.cpp:30-67
//   JSph::ConfigPosCellGpu          src/source/JSph.cpp:1949-1990
// with the same float/double types at every step (the cell hash later divides by double(float Scell)).
#include "../../include/b2s.h"
#include "b2s_internal.h"
#include <math.h>
#include <string.h>

static const double B2S_PI=3.14159265358979323846;

static unsigned calc_bits_value(unsigned v,unsigned minbits){
  minbits=(minbits<1||minbits>30? 1: minbits);
  unsigned n=minbits; for(;v>>n;n++);
  return n;
}
// JDsDcell::CalcCellDistribution (JDsDcell.cpp:41-57)
static bool calc_cell_distribution(const uint32_t nc[3],unsigned maxbits,unsigned out[3]){
  unsigned sx=calc_bits_value(nc[0],2), sy=calc_bits_value(nc[1],2), sz=calc_bits_value(nc[2],2);
  const unsigned smin=sx+sy+sz;
  if(smin<=maxbits){
    unsigned rest=maxbits-smin;
    while(rest){
      if(rest){ sx++; rest--; }
      if(rest){ sy++; rest--; }
      if(rest){ sz++; rest--; }
    }
  }
  out[0]=sx; out[1]=sy; out[2]=sz;
  return !(maxbits && sx+sy+sz>maxbits);
}

extern "C" int b2s_config_finalize(b2s_params* p){
  if(!p)return B2S_ERR_ARG;
  if(p->tkernel!=B2S_KERNEL_CUBIC && p->tkernel!=B2S_KERNEL_WENDLAND)return B2S_ERR_ARG;
  if(p->kernelh<=0 || p->dp<=0 || p->rhopzero<=0)return B2S_ERR_ARG;
  const bool sim2d=(p->simulate2d!=0);
  // ---- ConfigConstants1
  const double h=p->kernelh;                 // KernelH is float; promoted
  const double kh=h*2.0;                     // GetKernelFactor()==2 for both kernels
  p->kernelsize=float(kh);
  p->kernelsize2=p->kernelsize*p->kernelsize;
  if(sim2d){ p->awen=float(0.557/(h*h)); p->bwenh=float(-2.7852/(h*h*h*h)); }
  else     { p->awen=float(0.41778/(h*h*h)); p->bwenh=float(-2.08891/(h*h*h*h*h)); }
  {
    double a1,a2,aa;
    if(sim2d){ a1=10./(B2S_PI*7.); a2=a1/(h*h); aa=a1/(h*h*h); }
    else     { a1=1./B2S_PI;       a2=a1/(h*h*h); aa=a1/(h*h*h*h); }
    const double deltap=1./1.5;
    const double wdeltap=a2*(1.-1.5*deltap*deltap+0.75*deltap*deltap*deltap);
    p->cubic_od_wdeltap=float(1./wdeltap);
    p->cubic_a1=float(a1); p->cubic_a2=float(a2); p->cubic_aa=float(aa); p->cubic_a24=float(0.25*a2);
    p->cubic_c1=float(-3.*aa); p->cubic_d1=float(9.*aa/4.); p->cubic_c2=float(-3.*aa/4.);
  }
  p->cs0=sqrt(double(p->gamma)*double(p->cteb)/double(p->rhopzero));
  p->eta2=float((h*0.1)*(h*0.1));
  p->ovrhopzero=1.f/p->rhopzero;
  // ---- ConfigConstants2
  if(p->tvisco==B2S_VISCO_LAMINARSPS){
    const double dp_sps=(sim2d? sqrt(p->dp*p->dp*2.)/2.: sqrt(p->dp*p->dp*3.)/3.);
    p->spssmag=float((2./3.)*pow((0.12*dp_sps),2));
    p->spsblin=float((2./3.)*0.0066*dp_sps*dp_sps);
  }
  else{ p->spssmag=0; p->spsblin=0; }
  p->ddtkh=p->kernelsize*p->ddtvalue;
  p->ddtgz=float(double(p->rhopzero)*double(fabsf(p->gravity[2]))/double(p->cteb));
  if(p->coefdtmin<=0)p->coefdtmin=0.05f;     // default CoefDtMin (JSph.cpp:800), stored through float
  if(!p->dtini)p->dtini=double(p->kernelh)/p->cs0;
  if(!p->dtmin)p->dtmin=(double(p->kernelh)/p->cs0)*double(float(p->coefdtmin));
  // ---- map: Map_Pos == MapRealPos, widened by KernelSize on periodic axes for the duplicates (JSph.cpp:2184-2195)
  if(p->periactive<0 || p->periactive>7)return B2S_ERR_ARG;
  for(int k=0;k<3;k++){
    const bool peri=((p->periactive>>k)&1)!=0;
    p->maprealsize[k]=p->maprealposmax[k]-p->maprealposmin[k];
    p->map_posmin[k]=p->maprealposmin[k]-(peri? double(p->kernelsize): 0.0); p->map_posmax[k]=p->maprealposmax[k]+(peri? double(p->kernelsize): 0.0);
    p->map_size[k]=p->map_posmax[k]-p->map_posmin[k];
    if(!(p->map_size[k]>0))return B2S_ERR_ARG;
    double* inc=(k==0? p->perixinc: (k==1? p->periyinc: p->perizinc));
    if(peri)inc[k]=-p->maprealsize[k]; else inc[0]=inc[1]=inc[2]=0;
  }
  // ---- ConfigCellDivision
  p->scelldiv=(p->cellmode_half? 2: 1);
  p->scell=p->kernelsize/float(p->scelldiv);
  p->movlimit=p->scell*0.9f;
  for(int k=0;k<3;k++)p->map_cells[k]=unsigned(ceil(p->map_size[k]/p->scell));
  // ---- SelecDomain(0,Map_Cells)
  unsigned dc1[3];
  for(int k=0;k<3;k++){
    p->domcells[k]=p->map_cells[k];
    p->domposmin[k]=p->map_posmin[k];
    double dmax=p->map_posmin[k]+(p->domcells[k]*p->scell);
    if(dmax>p->map_posmax[k])dmax=p->map_posmax[k];
    p->domrealposmin[k]=p->domposmin[k]; p->domrealposmax[k]=dmax;
    if(p->domrealposmax[k]>p->maprealposmax[k])p->domrealposmax[k]=p->maprealposmax[k];
    if(p->domrealposmin[k]<p->maprealposmin[k])p->domrealposmin[k]=p->maprealposmin[k];
    dc1[k]=p->domcells[k]+1;
  }
  unsigned nb[3];
  if(!calc_cell_distribution(dc1,31,nb))return B2S_ERR_UNSUPPORTED;
  p->domcellcode=B2S_DCEL_GETCODE(nb[0],nb[1],nb[2]);
  // ---- ConfigPosCellGpu: 13/10/9 bits
  {
    const unsigned nx=1u<<13, ny=1u<<10, nz=1u<<9;
    float pcs=p->kernelsize; int nks=1;
    unsigned n[3];
    for(int k=0;k<3;k++)n[k]=unsigned(ceil(p->map_size[k]/pcs));
    while(n[0]>nx || n[1]>ny || n[2]>nz){
      nks=(nks==1? 2: nks+2);
      pcs=p->kernelsize*nks;
      for(int k=0;k<3;k++)n[k]=unsigned(ceil(p->map_size[k]/pcs));
    }
    p->poscellsize=pcs;
  }
  p->casenbound=p->casenp-p->casenfluid;
  p->casenpb=p->casenbound-p->casenfloat;
  p->use_floating=(p->casenfloat>0? 1: 0);
  return B2S_OK;
}

extern "C" const char* b2s_error_string(int code){
  switch(code){
    case B2S_OK: return "ok";
    case B2S_ERR_ARG: return "invalid argument";
    case B2S_ERR_NODEVICE: return "no CUDA device available (libb200sph has no CPU fallback)";
    case B2S_ERR_NOMEM: return "out of memory";
    case B2S_ERR_STATE: return "invalid state for this call";
    case B2S_ERR_UNSUPPORTED: return "configuration not supported";
    case B2S_ERR_DTINVALID: return "computed dt is NaN or infinity";
    case B2S_ERR_BOUNDOUT: return "boundary particle excluded from the domain";
    case B2S_ERR_IO: return "file i/o error";
    case B2S_ERR_NAME: return "unknown array name";
    default: return (code>0? "CUDA runtime error": "unknown error");
  }
}
