/* sph_oracle.c -- TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or called from the product
 * (libb200sph); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Plain-C restatement of the DualSPHysics v5.4 *CPU* implementation of the WCSPH time step, used as the
 * parity checker for the CUDA path.  Each function cites the reference code it follows
 * (paths relative to /root/reference/src/source).  Pinned against the real reference: tests/test_oracle_vs_ref.py
 * compares every function here with dumps produced by the compiled reference (oracle/_ref/dsph_ref_probe,
 * fixtures under tests/golden/: all 20 of them, i.e. every feature of the CUDA path -- cell divide, interaction with SPS and all DDT modes,
 * mDBC vel0 and mDBC2, Symplectic / Verlet, shifting, damping; periodic duplication, boundary motion and floating bodies live in
 * oracle/oracle.py).  Compiled WITHOUT -ffast-math so that it is bit-reproducible; the reference
 * build uses -ffast-math, so float results agree to rounding (tolerances in the tests), integers exactly.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint16_t typecode;

#define CODE_MASKSPECIAL 0xe000u
#define CODE_OUTIGNORE   0x4000u
#define CODE_OUTMOV      0x6000u
#define CODE_OUTPOS      0x8000u
#define CODE_OUTRHO      0xA000u
#define CODE_MASKTYPE    0x1800u
#define CODE_TYPE_FIXED    0x0000u
#define CODE_TYPE_FLOATING 0x1000u
#define CODE_TYPE_FLUID  0x1800u
#define CODE_MASKVALUE   0x07ffu
#define CODE_SPECIAL(c)  ((c)&CODE_MASKSPECIAL)
#define CODE_TYPE(c)     ((c)&CODE_MASKTYPE)
#define DCEL_CELLX(dcc,cel) ((cel)>>(((dcc)>>10)&31))
#define DCEL_CELLY(dcc,cel) (((cel)<<((dcc)>>25))>>(((dcc)>>5)&31))
#define DCEL_CELLZ(dcc,cel) (((cel)<<((dcc)&31))>>((dcc)&31))
#define DCEL_CELL(dcc,cx,cy,cz) (((cx)<<(((dcc)>>10)&31))|((cy)<<(((dcc)>>15)&31))|(cz))
#define ALMOSTZERO 1e-18f

/* Constants: the subset of StCteSph + JSph members the path reads (DualSphDef.h StCteSph, JSph.h). */
typedef struct orc_cte {
  int32_t tkernel, tvisco, tdensity, tboundary, tstep, use_floating, dt_allparticles, scelldiv;
  float kernelh, kernelsize2, awen, bwenh;
  float cub_a2, cub_a24, cub_c1, cub_d1, cub_c2, cub_od_wdeltap;
  float cteb, gamma, rhopzero, eta2, massfluid, massbound;
  float ddtkh, ddtgz, visco, viscoboundfactor, spssmag, spsblin;
  float scell, movlimit, rhopoutmin, rhopoutmax;
  float gravity[3];
  double cs0, cflnumber, dtmin;
  double domposmin[3], maprealposmin[3], maprealsize[3];
  uint32_t domcellcode, domcells[3];
  int32_t shiftmode;               /* TpShifting 0-3 (JSph.cpp:744-758) */
  float shiftcoef, shifttfs, dp;   /* ShiftCoef, ShiftTFS, float(Dp) */
  int32_t periactive, simulate2d;  /* periodic axes: bit0 X, bit1 Y, bit2 Z (JSph::PeriActive); Simulate2D (JSph.cpp:646) */
  double perinc[3][3];             /* PeriXinc, PeriYinc, PeriZinc (JSph.cpp:2186-2188) */
  double mapposmin[3], mapposmax[3]; /* Map_PosMin/Max: the real map widened by KernelSize on periodic axes (JSph.cpp:2191-2195) */
} orc_cte;

/* Result of a cell divide (JCellDivCpu members). */
typedef struct orc_div {
  uint32_t cellmin[3], cellmax[3], ncx, ncy, ncz, nsheet, nct, nctt;
  uint32_t box_boundignore, box_fluid, box_boundout, box_fluidout, box_boundoutignore, box_fluidoutignore;
  uint32_t np, npb, npbok, npb_ignore, npb_out, npf_out;
} orc_div;

/* ------------------------------------------------------------------------------------------------ */
/* JCellDivCpu::LimitsCellBound / LimitsCellFluid (JCellDivCpu.cpp:269-352)                          */
static void limits_cell(uint32_t n, int boundtype, const uint32_t* dcell, const typecode* code, uint32_t dcc,
                        uint32_t cmin[3], uint32_t cmax[3]){
  /* by particle TYPE over all n particles: equals the index ranges [0,npb) / [npb,np) of a plain run and also covers the
   * periodic duplicates appended behind them (CalcCellDomainBound/Fluid with the npb2/npf2 blocks, JCellDivCpuSingle.cpp:47-78) */
  cmin[0]=cmin[1]=cmin[2]=UINT32_MAX; cmax[0]=cmax[1]=cmax[2]=0;
  for(uint32_t p=0;p<n;p++){
    if(CODE_SPECIAL(code[p])<CODE_OUTIGNORE && (CODE_TYPE(code[p])<CODE_TYPE_FLOATING)==(boundtype!=0)){
      const uint32_t rc=dcell[p];
      const uint32_t c[3]={DCEL_CELLX(dcc,rc),DCEL_CELLY(dcc,rc),DCEL_CELLZ(dcc,rc)};
      for(int k=0;k<3;k++){ if(cmin[k]>c[k])cmin[k]=c[k]; if(cmax[k]<c[k])cmax[k]=c[k]; }
    }
  }
}

/* JCellDivCpuSingle::Divide (JCellDivCpuSingle.cpp:280-347): CalcCellDomain :47-78, MergeMapCellBoundFluid :90-107,
 * PrepareNct :117-135, PreSortFull :137-170, MakeSortFull :221-234.  Stable counting sort.
 * cellpart (sorted), sortpart, begincell[nctt] are outputs (begincell must hold 2*prod(domcells)+6 entries). */
int orc_celldiv(const orc_cte* c, uint32_t np, uint32_t npb, const uint32_t* dcell, const typecode* code,
                uint32_t* cellpart, uint32_t* sortpart, uint32_t* begincell, orc_div* d){
  uint32_t bmin[3],bmax[3],fmin[3],fmax[3];
  limits_cell(np,1,dcell,code,c->domcellcode,bmin,bmax);
  if(bmin[0]>bmax[0]){ for(int k=0;k<3;k++){ bmin[k]=c->domcells[k]; bmax[k]=0; } }
  limits_cell(np,0,dcell,code,c->domcellcode,fmin,fmax);
  if(fmin[0]>fmax[0]){ for(int k=0;k<3;k++){ fmin[k]=c->domcells[k]; fmax[k]=0; } }
  const uint32_t sd=(uint32_t)c->scelldiv;
  for(int k=0;k<3;k++){
    const uint32_t lo=(bmin[k]<fmin[k]? bmin[k]: fmin[k]), lo2=(fmin[k]>=sd? fmin[k]-sd: 0);
    d->cellmin[k]=(lo>lo2? lo: lo2);
    const uint32_t hi=(bmax[k]>fmax[k]? bmax[k]: fmax[k]), hi2=fmax[k]+sd;
    d->cellmax[k]=(hi<hi2? hi: hi2);
    if(d->cellmax[k]>=c->domcells[k])d->cellmax[k]=c->domcells[k]-1;
  }
  if(d->cellmin[0]>d->cellmax[0] || d->cellmin[1]>d->cellmax[1] || d->cellmin[2]>d->cellmax[2])
    for(int k=0;k<3;k++)d->cellmin[k]=d->cellmax[k]=0;
  d->ncx=d->cellmax[0]-d->cellmin[0]+1; d->ncy=d->cellmax[1]-d->cellmin[1]+1; d->ncz=d->cellmax[2]-d->cellmin[2]+1;
  d->nsheet=d->ncx*d->ncy; d->nct=d->nsheet*d->ncz; d->nctt=d->nct*2+5+1;
  d->box_boundignore=d->nct; d->box_fluid=d->nct+1; d->box_boundout=d->box_fluid+d->nct;
  d->box_fluidout=d->box_boundout+1; d->box_boundoutignore=d->box_fluidout+1; d->box_fluidoutignore=d->box_boundoutignore+1;
  uint32_t* partsincell=(uint32_t*)calloc(d->nctt,sizeof(uint32_t));
  uint32_t* cp=(uint32_t*)malloc(sizeof(uint32_t)*(np? np: 1));
  if(!partsincell||!cp){ free(partsincell); free(cp); return -1; }
  for(uint32_t p=0;p<np;p++){
    const uint32_t rcell=dcell[p];
    const uint32_t cx=DCEL_CELLX(c->domcellcode,rcell)-d->cellmin[0];
    const uint32_t cy=DCEL_CELLY(c->domcellcode,rcell)-d->cellmin[1];
    const uint32_t cz=DCEL_CELLZ(c->domcellcode,rcell)-d->cellmin[2];
    const uint32_t cellsort=cx+cy*d->ncx+cz*d->nsheet;
    const uint32_t codetype=CODE_TYPE(code[p]), codeout=CODE_SPECIAL(code[p]);
    uint32_t box;
    if(codetype<CODE_TYPE_FLOATING)
      box=(codeout<CODE_OUTIGNORE? ((cx<d->ncx && cy<d->ncy && cz<d->ncz)? cellsort: d->box_boundignore): (codeout==CODE_OUTIGNORE? d->box_boundoutignore: d->box_boundout));
    else
      box=(codeout<=CODE_OUTIGNORE? (codeout<CODE_OUTIGNORE? d->box_fluid+cellsort: d->box_fluidoutignore): (codetype==CODE_TYPE_FLOATING? d->box_boundout: d->box_fluidout));
    cp[p]=box; partsincell[box]++;
  }
  begincell[0]=0;
  for(uint32_t b=0;b<d->nctt-1;b++)begincell[b+1]=begincell[b]+partsincell[b];
  memset(partsincell,0,sizeof(uint32_t)*d->nctt);
  for(uint32_t p=0;p<np;p++){ const uint32_t b=cp[p]; sortpart[begincell[b]+partsincell[b]]=p; partsincell[b]++; }
  for(uint32_t p=0;p<np;p++)cellpart[p]=cp[sortpart[p]];
  #define CSIZE(b) (begincell[(b)+1]-begincell[(b)])
  d->npb_ignore=CSIZE(d->box_boundignore); d->npb_out=CSIZE(d->box_boundout); d->npf_out=CSIZE(d->box_fluidout);
  const uint32_t npboi=CSIZE(d->box_boundoutignore), npfoi=CSIZE(d->box_fluidoutignore);
  #undef CSIZE
  d->np=np-d->npb_out-d->npf_out-npboi-npfoi;
  d->npb=npb-npboi; d->npbok=d->npb-d->npb_ignore;
  free(partsincell); free(cp);
  return 0;
}

/* JCellDivCpu::SortArray: a2[p]=a[sortpart[p]] for one array of element size es. */
void orc_sort_array(uint32_t n, const uint32_t* sortpart, void* a, size_t es){
  char* src=(char*)a; char* tmp=(char*)malloc((size_t)(n? n: 1)*es);
  for(uint32_t p=0;p<n;p++)memcpy(tmp+(size_t)p*es,src+(size_t)sortpart[p]*es,es);
  memcpy(src,tmp,(size_t)n*es); free(tmp);
}

/* JSph::LoadDcellParticles (JSph.cpp:1794-1820) */
void orc_load_dcell(const orc_cte* c, uint32_t n, const typecode* code, const double* pos, uint32_t* dcell){
  for(uint32_t p=0;p<n;p++){
    if(CODE_SPECIAL(code[p])<CODE_OUTIGNORE){
      const double dx=pos[3*(size_t)p]-c->domposmin[0], dy=pos[3*(size_t)p+1]-c->domposmin[1], dz=pos[3*(size_t)p+2]-c->domposmin[2];
      const uint32_t cx=(uint32_t)(dx/c->scell), cy=(uint32_t)(dy/c->scell), cz=(uint32_t)(dz/c->scell);
      dcell[p]=DCEL_CELL(c->domcellcode,cx,cy,cz);
    }
    else dcell[p]=0xFFFFFFFFu;
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* fsph::GetKernel_Fac (FunSphKernel.h:53-65 cubic, :218-223 Wendland)                               */
static inline float kernel_fac(const orc_cte* c, float rr2){
  const float rad=sqrtf(rr2);
  const float qq=rad/c->kernelh;
  if(c->tkernel==2){ const float wqq1=1.f-0.5f*qq; return c->bwenh*wqq1*wqq1*wqq1; }
  if(rad>c->kernelh){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; return c->cub_c2*wqq2/rad; }
  { const float wqq2=qq*qq; return (c->cub_c1*qq+c->cub_d1*wqq2)/rad; }
}
/* fsph::GetKernelCubic_Tensil (FunSphKernel.h:88-97) */
static inline float cubic_tensil(const orc_cte* c, float rr2, float rhop1, float pressp1, float rhop2, float pressp2){
  const float rad=sqrtf(rr2), qq=rad/c->kernelh;
  float wab;
  if(rad>c->kernelh){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; wab=c->cub_a24*(wqq2*wqq1); }
  else{ const float wqq2=qq*qq; wab=c->cub_a2*(1.0f+(0.75f*qq-1.5f)*wqq2); }
  float fab=wab*c->cub_od_wdeltap; fab*=fab; fab*=fab;
  const float t1=(pressp1/(rhop1*rhop1))*(pressp1>0? 0.01f: -0.2f);
  const float t2=(pressp2/(rhop2*rhop2))*(pressp2>0? 0.01f: -0.2f);
  return fab*(t1+t2);
}
/* fsph::ComputePress (FunSphEos.h:37-39) */
void orc_press(const orc_cte* c, uint32_t n, const float* velrho, float* press){
  for(uint32_t p=0;p<n;p++)press[p]=c->cteb*(powf(velrho[4*(size_t)p+3]/c->rhopzero,c->gamma)-1.0f);
}

/* nsearch::Init + ParticleRange (JCellSearch_inline.h): neighbour rows of the particle's cell. */
typedef struct { int cx,cy,cz,xini,xfin,yini,yfin,zini,zfin; } ngs_t;
static inline ngs_t ngs_init(const orc_cte* c, const orc_div* d, uint32_t dcell){
  ngs_t r;
  r.cx=(int)DCEL_CELLX(c->domcellcode,dcell)-(int)d->cellmin[0];
  r.cy=(int)DCEL_CELLY(c->domcellcode,dcell)-(int)d->cellmin[1];
  r.cz=(int)DCEL_CELLZ(c->domcellcode,dcell)-(int)d->cellmin[2];
  const int s=c->scelldiv, ncx=(int)d->ncx, ncy=(int)d->ncy, ncz=(int)d->ncz;
  #define MIN_(a,b) ((a)<(b)?(a):(b))
  r.xini=r.cx-MIN_(r.cx,s); r.xfin=r.cx+MIN_(ncx-r.cx-1,s)+1;
  r.yini=r.cy-MIN_(r.cy,s); r.yfin=r.cy+MIN_(ncy-r.cy-1,s)+1;
  r.zini=r.cz-MIN_(r.cz,s); r.zfin=r.cz+MIN_(ncz-r.cz-1,s)+1;
  #undef MIN_
  return r;
}

/* JSphCpu::InteractionForcesFluid (JSphCpu.cpp:736-1125), one pass (boundp2=0 fluid/floating, 1 boundary).
 * delta[] follows the CPU build's DDTArray=true path (JSph.cpp:927; JSphCpu.cpp:1088-1091). */
static void forces_fluid_pass(const orc_cte* c, const orc_div* d, int boundp2, float visco, const uint32_t* begincell,
  const uint32_t* dcell, const double* pos, const float* velrho, const typecode* code, const float* press,
  const float* tau, float* two_strain, const float* ftomassp, float* viscdt_out, float* ar, float* ace, float* delta,
  const uint8_t* boundmode, const float* tangenvel)
{
  float viscdt=*viscdt_out;
  const int m2=(boundp2 && boundmode!=NULL);   /* mDBC2: boundary neighbours use boundmode / tangenvel (JSphCpu.cpp:851-853,990-995) */
  const uint32_t* beg=begincell+(boundp2? 0: d->box_fluid);
  #pragma omp parallel for schedule(guided) reduction(max:viscdt)
  for(int64_t p1=(int64_t)d->npb;p1<(int64_t)d->np;p1++){
    float visc=0,arp1=0,deltap1=0, ax=0,ay=0,az=0;
    float st[6]={0,0,0,0,0,0};
    int ftp1=0;
    if(c->use_floating){ ftp1=(CODE_TYPE(code[p1])==CODE_TYPE_FLOATING); if(ftp1 && c->tdensity!=0)deltap1=FLT_MAX; }
    const double px=pos[3*p1],py=pos[3*p1+1],pz=pos[3*p1+2];
    const float v1x=velrho[4*p1],v1y=velrho[4*p1+1],v1z=velrho[4*p1+2],rhop1=velrho[4*p1+3];
    const float pressp1=press[p1];
    const float* taup1=(tau? tau+6*p1: NULL);
    const ngs_t r=ngs_init(c,d,dcell[p1]);
    for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
      const uint32_t pini=beg[r.xini+y*(int)d->ncx+z*(int)d->nsheet], pfin=beg[r.xfin+y*(int)d->ncx+z*(int)d->nsheet];
      for(uint32_t p2=pini;p2<pfin;p2++){
        const float drx=(float)(px-pos[3*(size_t)p2]), dry=(float)(py-pos[3*(size_t)p2+1]), drz=(float)(pz-pos[3*(size_t)p2+2]);
        const float rr2=drx*drx+dry*dry+drz*drz;
        if(rr2<=c->kernelsize2 && rr2>=ALMOSTZERO){
          const float fac=kernel_fac(c,rr2);
          const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
          float massp2=(boundp2? c->massbound: c->massfluid);
          int ftp2=0;
          if(c->use_floating){
            ftp2=(CODE_TYPE(code[p2])==CODE_TYPE_FLOATING);
            if(ftp2)massp2=ftomassp[code[p2]&CODE_MASKVALUE];
            if(ftp2 && c->tdensity==1 && massp2<=(c->massfluid*1.2f))deltap1=FLT_MAX;   /* DELTA_HEAVYFLOATING */
          }
          if(m2 && !ftp2 && boundmode[p2]==2)massp2=0;   /* BMODE_MDBC2OFF */
          const float v2x=velrho[4*(size_t)p2],v2y=velrho[4*(size_t)p2+1],v2z=velrho[4*(size_t)p2+2],rhop2=velrho[4*(size_t)p2+3];
          { /* momentum */
            const float prs=(pressp1+press[p2])/(rhop1*rhop2)+(c->tkernel==1? cubic_tensil(c,rr2,rhop1,pressp1,rhop2,press[p2]): 0);
            const float p_vpm=-prs*massp2;
            ax+=p_vpm*frx; ay+=p_vpm*fry; az+=p_vpm*frz;
          }
          float dvx=v1x-v2x,dvy=v1y-v2y,dvz=v1z-v2z;
          arp1+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhop1/rhop2);
          const float cbar=(float)c->cs0;
          if(c->tdensity==1 && deltap1!=FLT_MAX){
            const float rhop1over2=rhop1/rhop2;
            const float visc_densi=c->ddtkh*cbar*(rhop1over2-1.f)/(rr2+c->eta2);
            const float dot3=(drx*frx+dry*fry+drz*frz);
            const float delta_=visc_densi*dot3*massp2;
            deltap1=(boundp2 && c->tboundary==1? FLT_MAX: deltap1+delta_);
          }
          if((c->tdensity==2 || (c->tdensity==3 && !boundp2)) && deltap1!=FLT_MAX && !ftp2){
            const float rh=1.f+c->ddtgz*drz;
            const float drho=c->rhopzero*powf(rh,1.f/c->gamma)-c->rhopzero;
            const float visc_densi=c->ddtkh*cbar*((rhop2-rhop1)-drho)/(rr2+c->eta2);
            const float dot3=(drx*frx+dry*fry+drz*frz);
            const float delta_=visc_densi*dot3*massp2/rhop2;
            deltap1=(boundp2? FLT_MAX: deltap1-delta_);
          }
          { /* viscosity */
            if(m2 && !ftp2){ const float* tv=tangenvel+3*(size_t)p2; dvx=v1x-tv[0]; dvy=v1y-tv[1]; dvz=v1z-tv[2]; }
            const float dot=drx*dvx+dry*dvy+drz*dvz;
            const float dot_rr2=dot/(rr2+c->eta2);
            if(dot_rr2>visc)visc=dot_rr2;
            if(c->tvisco==1){
              if(dot<0){
                const float amubar=c->kernelh*dot_rr2;
                const float robar=(rhop1+rhop2)*0.5f;
                const float pi_visc=(-visco*cbar*amubar/robar)*massp2;
                ax-=pi_visc*frx; ay-=pi_visc*fry; az-=pi_visc*frz;
              }
            }
            else{
              const float robar2=(rhop1+rhop2);
              const float temp=4.f*visco/((rr2+c->eta2)*robar2);
              const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
              ax+=vtemp*dvx; ay+=vtemp*dvy; az+=vtemp*dvz;
            }
            if(c->tvisco==2){
              float sxx=taup1[0],sxy=taup1[1],sxz=taup1[2],syy=taup1[3],syz=taup1[4],szz=taup1[5];
              if(!boundp2 && !ftp2){ const float* t2=tau+6*(size_t)p2; sxx+=t2[0]; sxy+=t2[1]; sxz+=t2[2]; syy+=t2[3]; syz+=t2[4]; szz+=t2[5]; }
              ax+=massp2*(sxx*frx+sxy*fry+sxz*frz);
              ay+=massp2*(sxy*frx+syy*fry+syz*frz);
              az+=massp2*(sxz*frx+syz*fry+szz*frz);
              if(!ftp1){
                const float volp2=-massp2/rhop2;
                float dv=dvx*volp2; st[0]+=dv*frx; st[1]+=dv*fry; st[2]+=dv*frz;
                      dv=dvy*volp2; st[1]+=dv*frx; st[3]+=dv*fry; st[4]+=dv*frz;
                      dv=dvz*volp2; st[2]+=dv*frx; st[4]+=dv*fry; st[5]+=dv*frz;
              }
            }
          }
        }
      }
    }
    if(c->shiftmode||arp1||ax||ay||az||visc){   /* `if(shift||arp1||...)`, JSphCpu.cpp:1084 */
      if(c->tdensity!=0)delta[p1]=(delta[p1]==FLT_MAX || deltap1==FLT_MAX? FLT_MAX: delta[p1]+deltap1);
      ar[p1]+=arp1;
      ace[3*p1]+=ax; ace[3*p1+1]+=ay; ace[3*p1+2]+=az;
      if(visc>viscdt)viscdt=visc;
      if(c->tvisco==2)for(int k=0;k<6;k++)two_strain[6*p1+k]+=st[k];
    }
  }
  *viscdt_out=viscdt;
}

/* JSphCpu::InteractionForcesBound (JSphCpu.cpp:607-700) */
static void forces_bound(const orc_cte* c, const orc_div* d, const uint32_t* begincell, const uint32_t* dcell, const double* pos,
  const float* velrho, const typecode* code, const float* ftomassp, float* viscdt_out, float* ar)
{
  float viscdt=*viscdt_out;
  const uint32_t* beg=begincell+d->box_fluid;
  #pragma omp parallel for schedule(guided) reduction(max:viscdt)
  for(int64_t p1=0;p1<(int64_t)d->npbok;p1++){
    float visc=0,arp1=0;
    const double px=pos[3*p1],py=pos[3*p1+1],pz=pos[3*p1+2];
    const float v1x=velrho[4*p1],v1y=velrho[4*p1+1],v1z=velrho[4*p1+2],rhop1=velrho[4*p1+3];
    const ngs_t r=ngs_init(c,d,dcell[p1]);
    for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
      const uint32_t pini=beg[r.xini+y*(int)d->ncx+z*(int)d->nsheet], pfin=beg[r.xfin+y*(int)d->ncx+z*(int)d->nsheet];
      for(uint32_t p2=pini;p2<pfin;p2++){
        const float drx=(float)(px-pos[3*(size_t)p2]), dry=(float)(py-pos[3*(size_t)p2+1]), drz=(float)(pz-pos[3*(size_t)p2+2]);
        const float rr2=drx*drx+dry*dry+drz*drz;
        if(rr2<=c->kernelsize2 && rr2>=ALMOSTZERO){
          const float fac=kernel_fac(c,rr2);
          const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
          float massp2=c->massfluid;
          if(c->use_floating && CODE_TYPE(code[p2])==CODE_TYPE_FLOATING)massp2=ftomassp[code[p2]&CODE_MASKVALUE];
          const float dvx=v1x-velrho[4*(size_t)p2],dvy=v1y-velrho[4*(size_t)p2+1],dvz=v1z-velrho[4*(size_t)p2+2];
          arp1+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhop1/velrho[4*(size_t)p2+3]);
          const float dot=drx*dvx+dry*dvy+drz*dvz;
          const float dot_rr2=dot/(rr2+c->eta2);
          if(dot_rr2>visc)visc=dot_rr2;
        }
      }
    }
    if(arp1||visc){ ar[p1]+=arp1; if(visc>viscdt)viscdt=visc; }
  }
  *viscdt_out=viscdt;
}

/* JSphCpu::ComputeSpsTau (JSphCpu.cpp:1279-1307) */
static void sps_tau(const orc_cte* c, uint32_t n, uint32_t pini, const float* velrho, const float* s2, float* tau){
  for(uint32_t p=pini;p<pini+n;p++){
    const float* t=s2+6*(size_t)p;
    const float pow1=t[0]*t[0]+t[3]*t[3]+t[5]*t[5];
    const float prr=t[1]*t[1]+t[2]*t[2]+t[4]*t[4]+pow1+pow1;
    const float visc_sps=c->spssmag*sqrtf(prr);
    const float div_u=t[0]+t[3]+t[5];
    const float sps_k=visc_sps*div_u;
    const float sps_blin=c->spsblin*prr;
    const float sumsps=-(sps_k+sps_blin);
    const float twovisc_sps=(visc_sps+visc_sps);
    const float one_rho=1.0f/velrho[4*(size_t)p+3];
    float* o=tau+6*(size_t)p;
    o[0]=one_rho*(twovisc_sps*t[0]+sumsps); o[1]=one_rho*(visc_sps*t[1]); o[2]=one_rho*(visc_sps*t[2]);
    o[3]=one_rho*(twovisc_sps*t[3]+sumsps); o[4]=one_rho*(visc_sps*t[4]); o[5]=one_rho*(twovisc_sps*t[5]+sumsps);
  }
}

/* PreInteraction_Forces + Interaction_Forces (JSphCpu.cpp:488-557, JSphCpuSingle.cpp:685-739, JSphCpu.cpp:1318-1375).
 * Outputs: ace[3np], ar[np], press[np], maxima {velmax, acemax, viscdtmax}. tau (in/out) may be NULL. */
int orc_interaction(const orc_cte* c, const orc_div* d, const uint32_t* begincell, const uint32_t* dcell, const double* pos,
  const float* velrho, const typecode* code, float* tau, const float* ftomassp,
  float* ace, float* ar, float* press, double* velmax, double* acemax, float* viscdtmax,
  const uint8_t* boundmode, const float* tangenvel)
{
  const uint32_t np=d->np, npb=d->npb;
  memset(ace,0,sizeof(float)*3*(size_t)np); memset(ar,0,sizeof(float)*(size_t)np);
  float* delta=(float*)calloc(np? np: 1,sizeof(float));
  float* s2=(c->tvisco==2? (float*)calloc((size_t)(np? np: 1)*6,sizeof(float)): NULL);
  orc_press(c,np,velrho,press);
  { /* CalcVelMaxSeq (JSphCpu.cpp:562-570) */
    const uint32_t pini=(c->dt_allparticles? 0: npb); float vm=0;
    for(uint32_t p=pini;p<np;p++){ const float* v=velrho+4*(size_t)p; const float v2=v[0]*v[0]+v[1]*v[1]+v[2]*v[2]; if(v2>vm)vm=v2; }
    *velmax=sqrtf(vm);
  }
  float viscdt=0;
  if(np>npb){
    forces_fluid_pass(c,d,0,c->visco,begincell,dcell,pos,velrho,code,press,tau,s2,ftomassp,&viscdt,ar,ace,delta,boundmode,tangenvel);
    forces_fluid_pass(c,d,1,c->visco*c->viscoboundfactor,begincell,dcell,pos,velrho,code,press,tau,s2,ftomassp,&viscdt,ar,ace,delta,boundmode,tangenvel);
    if(c->tvisco==2)sps_tau(c,np-npb,npb,velrho,s2,tau);
  }
  if(d->npbok)forces_bound(c,d,begincell,dcell,pos,velrho,code,ftomassp,&viscdt,ar);
  if(c->simulate2d)for(uint32_t p=npb;p<np;p++)ace[3*(size_t)p+1]=0.f;                   /* JSphCpuSingle.cpp:717-724 */
  if(c->tdensity!=0)for(uint32_t p=npb;p<np;p++)if(delta[p]!=FLT_MAX)ar[p]+=delta[p];   /* JSphCpuSingle.cpp:723-731 */
  { /* ComputeAceMaxSeq (JSphCpuSingle.cpp:793-805) */
    float am=0; for(uint32_t p=npb;p<np;p++){ if(c->periactive && CODE_SPECIAL(code[p])!=0)continue;   /* ComputeAceMaxSeq<checkcode> with periodic */
      const float* a=ace+3*(size_t)p; const float a2=a[0]*a[0]+a[1]*a[1]+a[2]*a[2]; if(a2>am)am=a2; }
    *acemax=sqrt((double)am);
  }
  *viscdtmax=viscdt;
  free(delta); free(s2);
  return 0;
}

/* JSphCpu::DtVariable (JSphCpu.cpp:2031-2076) */
double orc_dtvariable(const orc_cte* c, double velmax, double acemax, float viscdtmax){
  const double dt1=(acemax? (sqrt((double)c->kernelh/acemax)): DBL_MAX);
  const double mx=(c->cs0>velmax*10.? c->cs0: velmax*10.);
  const double dt2=(double)c->kernelh/(mx+(double)c->kernelh*viscdtmax);
  double dt=c->cflnumber*(dt1<dt2? dt1: dt2);
  if(dt<c->dtmin)dt=c->dtmin;
  return dt;
}

/* JSphCpu::UpdatePos (JSphCpu.cpp:1458-1510), no periodic boundaries. */
/* JSph::UpdatePeriodicPos / the periodic branch of JSphCpu::UpdatePos (JSph.cpp:2079-2103, JSphCpu.cpp:1477-1495): returns out. */
static int periodic_wrap(const orc_cte* c, double* rx,double* ry,double* rz){
  double dx=*rx-c->maprealposmin[0], dy=*ry-c->maprealposmin[1], dz=*rz-c->maprealposmin[2];
  int out=(dx!=dx || dy!=dy || dz!=dz || dx<0 || dy<0 || dz<0 || dx>=c->maprealsize[0] || dy>=c->maprealsize[1] || dz>=c->maprealsize[2]);
  if(c->periactive && out){
    const int px=(c->periactive&1), py=(c->periactive&2), pz=(c->periactive&4);
    if(px){
      if(dx<0)                 { dx-=c->perinc[0][0]; dy-=c->perinc[0][1]; dz-=c->perinc[0][2]; }
      if(dx>=c->maprealsize[0]){ dx+=c->perinc[0][0]; dy+=c->perinc[0][1]; dz+=c->perinc[0][2]; }
    }
    if(py){
      if(dy<0)                 { dx-=c->perinc[1][0]; dy-=c->perinc[1][1]; dz-=c->perinc[1][2]; }
      if(dy>=c->maprealsize[1]){ dx+=c->perinc[1][0]; dy+=c->perinc[1][1]; dz+=c->perinc[1][2]; }
    }
    if(pz){
      if(dz<0)                 { dx-=c->perinc[2][0]; dy-=c->perinc[2][1]; dz-=c->perinc[2][2]; }
      if(dz>=c->maprealsize[2]){ dx+=c->perinc[2][0]; dy+=c->perinc[2][1]; dz+=c->perinc[2][2]; }
    }
    out=((!px && (dx<0 || dx>=c->maprealsize[0])) || (!py && (dy<0 || dy>=c->maprealsize[1])) || (!pz && (dz<0 || dz>=c->maprealsize[2])));
    *rx=dx+c->maprealposmin[0]; *ry=dy+c->maprealposmin[1]; *rz=dz+c->maprealposmin[2];
  }
  return out;
}

/* JSphCpu::UpdatePos (JSphCpu.cpp:1458-1516). */
static void update_pos(const orc_cte* c, double rx,double ry,double rz,double movx,double movy,double movz,int outrho,
  uint32_t p,double* pos,uint32_t* cell,typecode* code)
{
  const int outmov=(fabsf((float)movx)>c->movlimit || fabsf((float)movy)>c->movlimit || fabsf((float)movz)>c->movlimit);
  rx+=movx; ry+=movy; rz+=movz;
  const int out=periodic_wrap(c,&rx,&ry,&rz);
  pos[3*(size_t)p]=rx; pos[3*(size_t)p+1]=ry; pos[3*(size_t)p+2]=rz;
  if(outrho || outmov || out){
    typecode rcode=code[p];
    rcode=(typecode)((rcode&(~CODE_MASKSPECIAL))|(out? CODE_OUTPOS: (outrho? CODE_OUTRHO: CODE_OUTMOV)));
    code[p]=rcode; cell[p]=0xFFFFFFFFu;
  }
  else{
    /* relative to DomPosMin; without periodic axes that is MapRealPosMin (JSphCpu.cpp:1508-1514) */
    const double dx=rx-c->domposmin[0], dy=ry-c->domposmin[1], dz=rz-c->domposmin[2];
    const uint32_t cx=(uint32_t)(dx/c->scell), cy=(uint32_t)(dy/c->scell), cz=(uint32_t)(dz/c->scell);
    cell[p]=DCEL_CELL(c->domcellcode,cx,cy,cz);
  }
}

/* JSphCpu::MoveLinBound / MoveMatBound position part (JSphCpu.cpp:2105-2173): UpdatePos with the displacement of each listed
 * particle (outrho=false); the caller sets the velocities. */
void orc_move_particles(const orc_cte* c, uint32_t n, const uint32_t* pidx, const double* mov, double* pos, uint32_t* dcell, typecode* code){
  for(uint32_t k=0;k<n;k++){
    const uint32_t p=pidx[k];
    update_pos(c,pos[3*(size_t)p],pos[3*(size_t)p+1],pos[3*(size_t)p+2],mov[3*(size_t)k],mov[3*(size_t)k+1],mov[3*(size_t)k+2],0,p,pos,dcell,code);
  }
}

/* JSphCpu::ComputeSymplecticPre (JSphCpu.cpp:1687-1850): (pospre,velrhopre) -> (pos,velrho) */
void orc_symplectic_pre(const orc_cte* c, uint32_t np, uint32_t npb, double dt, const double* pospre, const float* velrhopre,
  const float* ar, const float* ace, double* pos, float* velrho, uint32_t* dcell, typecode* code, const uint8_t* boundmode)
{
  const double dt05=dt*.5;
  for(uint32_t p=0;p<npb;p++){
    float w=velrhopre[4*(size_t)p+3];
    if(!boundmode || boundmode[p]<1){   /* DBC or mDBC vel0; mDBC2 keeps the extrapolated density (JSphCpu.cpp:1717) */
      w=(float)((double)w+dt05*ar[p]);
      w=(w<c->rhopzero? c->rhopzero: w);
    }
    velrho[4*(size_t)p]=velrhopre[4*(size_t)p]; velrho[4*(size_t)p+1]=velrhopre[4*(size_t)p+1]; velrho[4*(size_t)p+2]=velrhopre[4*(size_t)p+2]; velrho[4*(size_t)p+3]=w;
    memcpy(pos+3*(size_t)p,pospre+3*(size_t)p,24);
  }
  for(uint32_t p=npb;p<np;p++){
    const float* vp=velrhopre+4*(size_t)p;
    const float rhonew=(float)((double)vp[3]+dt05*ar[p]);
    if(!c->use_floating || CODE_TYPE(code[p])==CODE_TYPE_FLUID){
      const double dx=(double)vp[0]*dt05, dy=(double)vp[1]*dt05, dz=(double)vp[2]*dt05;
      const int outrho=(rhonew<c->rhopoutmin || rhonew>c->rhopoutmax);
      velrho[4*(size_t)p]  =(float)((double)vp[0]+((double)ace[3*(size_t)p]  +c->gravity[0])*dt05);
      velrho[4*(size_t)p+1]=(float)((double)vp[1]+((double)ace[3*(size_t)p+1]+c->gravity[1])*dt05);
      velrho[4*(size_t)p+2]=(float)((double)vp[2]+((double)ace[3*(size_t)p+2]+c->gravity[2])*dt05);
      velrho[4*(size_t)p+3]=rhonew;
      update_pos(c,pospre[3*(size_t)p],pospre[3*(size_t)p+1],pospre[3*(size_t)p+2],dx,dy,dz,outrho,p,pos,dcell,code);
    }
    else{
      memcpy(velrho+4*(size_t)p,vp,16); velrho[4*(size_t)p+3]=(rhonew<c->rhopzero? c->rhopzero: rhonew);
      memcpy(pos+3*(size_t)p,pospre+3*(size_t)p,24);
    }
  }
}

/* JSphCpu::ComputeSymplecticCorr (JSphCpu.cpp:1856-2022): velrho holds the predicted state on entry. */
void orc_symplectic_corr(const orc_cte* c, uint32_t np, uint32_t npb, double dt, const double* pospre, const float* velrhopre,
  const float* ar, const float* ace, double* pos, float* velrho, uint32_t* dcell, typecode* code, const float* shiftpos, const uint8_t* boundmode)
{
  const double dt05=dt*.5;
  for(uint32_t p=0;p<np;p++){
    const float* vp=velrhopre+4*(size_t)p;
    if(p<npb && boundmode && boundmode[p]>=1){ memcpy(velrho+4*(size_t)p,vp,16); continue; }   /* mDBC2: velrho of the start of the step (JSphCpu.cpp:1869-1875) */
    const double epsilon_rdot=(-(double)ar[p]/(double)velrho[4*(size_t)p+3])*dt;
    const float rhonew=(float)((double)vp[3]*(2.-epsilon_rdot)/(2.+epsilon_rdot));
    if(p<npb){ velrho[4*(size_t)p]=velrho[4*(size_t)p+1]=velrho[4*(size_t)p+2]=0; velrho[4*(size_t)p+3]=(rhonew<c->rhopzero? c->rhopzero: rhonew); continue; }
    if(!c->use_floating || CODE_TYPE(code[p])==CODE_TYPE_FLUID){
      const float nx=(float)((double)vp[0]+((double)ace[3*(size_t)p]  +c->gravity[0])*dt);
      const float ny=(float)((double)vp[1]+((double)ace[3*(size_t)p+1]+c->gravity[1])*dt);
      const float nz=(float)((double)vp[2]+((double)ace[3*(size_t)p+2]+c->gravity[2])*dt);
      double dx=((double)vp[0]+(double)nx)*dt05, dy=((double)vp[1]+(double)ny)*dt05, dz=((double)vp[2]+(double)nz)*dt05;
      if(shiftpos){ dx+=(double)shiftpos[4*(size_t)p]; dy+=(double)shiftpos[4*(size_t)p+1]; dz+=(double)shiftpos[4*(size_t)p+2]; }   /* JSphCpu.cpp:1936-1940 */
      const int outrho=(rhonew<c->rhopoutmin || rhonew>c->rhopoutmax);
      velrho[4*(size_t)p]=nx; velrho[4*(size_t)p+1]=ny; velrho[4*(size_t)p+2]=nz; velrho[4*(size_t)p+3]=rhonew;
      update_pos(c,pospre[3*(size_t)p],pospre[3*(size_t)p+1],pospre[3*(size_t)p+2],dx,dy,dz,outrho,p,pos,dcell,code);
    }
    else{
      memcpy(velrho+4*(size_t)p,vp,16); velrho[4*(size_t)p+3]=(rhonew<c->rhopzero? c->rhopzero: rhonew);
      memcpy(pos+3*(size_t)p,pospre+3*(size_t)p,24);
    }
  }
}

/* JSphCpu::ComputeVerlet (JSphCpu.cpp:1523-1680): new values -> velrhonew; velrho1=v(n), velrho2=v(n-1) or v(n) on Euler steps. */
void orc_verlet(const orc_cte* c, uint32_t np, uint32_t npb, double dt, int euler, const float* velrho1, const float* velrho2,
  const float* ar, const float* ace, double* pos, uint32_t* dcell, typecode* code, float* velrhonew, const float* shiftpos, const uint8_t* boundmode)
{
  const double dt2=(euler? dt: dt+dt), dt205=0.5*dt*dt;
  for(uint32_t p=0;p<np;p++){
    const float* v1=velrho1+4*(size_t)p; const float* v2=velrho2+4*(size_t)p;
    if(p<npb && boundmode && boundmode[p]>=1){ memcpy(velrhonew+4*(size_t)p,v2,16); continue; }   /* ComputeVelrhoBound (JSphCpu.cpp:1616) */
    const float rhonew=(float)((double)v2[3]+dt2*ar[p]);
    float o[4];
    if(p<npb){ o[0]=o[1]=o[2]=0; o[3]=(rhonew<c->rhopzero? c->rhopzero: rhonew); }
    else if(!c->use_floating || CODE_TYPE(code[p])==CODE_TYPE_FLUID){
      const double agx=(double)ace[3*(size_t)p]+(double)c->gravity[0], agy=(double)ace[3*(size_t)p+1]+(double)c->gravity[1], agz=(double)ace[3*(size_t)p+2]+(double)c->gravity[2];
      double dx=(double)v1[0]*dt+agx*dt205, dy=(double)v1[1]*dt+agy*dt205, dz=(double)v1[2]*dt+agz*dt205;
      if(shiftpos){ dx+=(double)shiftpos[4*(size_t)p]; dy+=(double)shiftpos[4*(size_t)p+1]; dz+=(double)shiftpos[4*(size_t)p+2]; }   /* JSphCpu.cpp:1543-1547 */
      const int outrho=(rhonew<c->rhopoutmin || rhonew>c->rhopoutmax);
      o[0]=(float)((double)v2[0]+agx*dt2); o[1]=(float)((double)v2[1]+agy*dt2); o[2]=(float)((double)v2[2]+agz*dt2); o[3]=rhonew;
      update_pos(c,pos[3*(size_t)p],pos[3*(size_t)p+1],pos[3*(size_t)p+2],dx,dy,dz,outrho,p,pos,dcell,code);
    }
    else{ o[0]=v1[0]; o[1]=v1[1]; o[2]=v1[2]; o[3]=(rhonew<c->rhopzero? c->rhopzero: rhonew); }
    memcpy(velrhonew+4*(size_t)p,o,16);
  }
}

/* Standard shifting without zones: the shiftposfs sums of the pair loop (JSphCpu.cpp:767-779,848,942-950: fluid neighbours,
 * then boundary neighbours, one running sum) followed by JSphShifting::RunCpu (JSphShifting.cpp:390-423).  shiftpos = 4 floats per
 * particle: displacement xyz of the coming position update, w = the free-surface divergence. */
void orc_shifting(const orc_cte* c, const orc_div* d, const uint32_t* begincell, const uint32_t* dcell, const double* pos,
  const float* velrho, const typecode* code, const float* ftomassp, double dt, float* shiftpos)
{
  const double coefumagn=dt*(double)c->shiftcoef*(double)c->kernelh;
  const double coeftfs=(c->simulate2d? 2.0: 3.0)-(double)c->shifttfs;   /* JSphShifting.cpp:400 */
  const float maxdist=(float)((double)c->dp*0.1);
  memset(shiftpos,0,sizeof(float)*4*(size_t)d->np);
  #pragma omp parallel for schedule(guided)
  for(int64_t p1=(int64_t)d->npb;p1<(int64_t)d->np;p1++){
    float sx=0,sy=0,sz=0,sw=0;
    if(c->use_floating && CODE_TYPE(code[p1])==CODE_TYPE_FLOATING)sx=FLT_MAX;
    const double px=pos[3*p1],py=pos[3*p1+1],pz=pos[3*p1+2];
    const ngs_t r=ngs_init(c,d,dcell[p1]);
    for(int boundp2=0;boundp2<2 && sx!=FLT_MAX;boundp2++){
      const uint32_t* beg=begincell+(boundp2? 0: d->box_fluid);
      for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
        const uint32_t pini=beg[r.xini+y*(int)d->ncx+z*(int)d->nsheet], pfin=beg[r.xfin+y*(int)d->ncx+z*(int)d->nsheet];
        for(uint32_t p2=pini;p2<pfin;p2++){
          const float drx=(float)(px-pos[3*(size_t)p2]), dry=(float)(py-pos[3*(size_t)p2+1]), drz=(float)(pz-pos[3*(size_t)p2+2]);
          const float rr2=drx*drx+dry*dry+drz*drz;
          if(rr2<=c->kernelsize2 && rr2>=ALMOSTZERO){
            const float fac=kernel_fac(c,rr2);
            const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
            float massp2=(boundp2? c->massbound: c->massfluid);
            if(c->use_floating && CODE_TYPE(code[p2])==CODE_TYPE_FLOATING){
              massp2=ftomassp[code[p2]&CODE_MASKVALUE];
              if(c->shiftmode==1)sx=FLT_MAX;
            }
            if(sx!=FLT_MAX){
              const float massrho=massp2/velrho[4*(size_t)p2+3];
              const int noshift=(boundp2 && (c->shiftmode==1 || (c->shiftmode==2 && CODE_TYPE(code[p2])==CODE_TYPE_FIXED)));
              sx=(noshift? FLT_MAX: sx+massrho*frx);
              sy+=massrho*fry; sz+=massrho*frz;
              sw-=massrho*(drx*frx+dry*fry+drz*frz);
            }
          }
        }
      }
    }
    float* o=shiftpos+4*(size_t)p1;
    if(sx!=FLT_MAX){
      const double vx=velrho[4*p1],vy=velrho[4*p1+1],vz=velrho[4*p1+2];
      double umagn=coefumagn*sqrt(vx*vx+vy*vy+vz*vz);
      if(c->shifttfs){
        if(sw<c->shifttfs)umagn=0;
        else umagn*=((double)sw-(double)c->shifttfs)/coeftfs;
      }
      const float dx=(float)((double)sx*umagn), dy=(float)((double)sy*umagn), dz=(float)((double)sz*umagn);
      o[0]=(fabsf(dx)<maxdist? dx: (dx>=0? maxdist: -maxdist));
      o[1]=(fabsf(dy)<maxdist? dy: (dy>=0? maxdist: -maxdist));
      o[2]=(fabsf(dz)<maxdist? dz: (dz>=0? maxdist: -maxdist));
      o[3]=sw;
    }
  }
}

/* Damping zones (JDsDampingOp_Plane/_Box::ComputeDampingCpu, JDsDamping.cpp:185-233,446-492), zones applied in list order.
 * One zone = 64 doubles: [0] type (0 plane, 1 box), [1] overlimit, [2] redumax, [3..5] factorxyz,
 *   plane: [6..9] plane a,b,c,d, [10] dist, [11] usedomain, [12] zmin, [13] zmax, [14..29] four lateral planes;
 *   box:   [6..8] limmin1, [9..11] limmin2, [12..14] limmax1, [15..17] limmax2, [18..20] limover1, [21..23] limover2,
 *          [24..26] boxsize1, [27..29] boxsize2   (all already derived like JDsDampingOp_Box::ReadXml, JDsDamping.cpp:345-391). */
void orc_damping(uint32_t np, uint32_t npb, double dt, uint32_t nzones, const double* zones, const double* pos, const typecode* code, float* velrho)
{
  for(uint32_t k=0;k<nzones;k++){
    const double* z=zones+64*(size_t)k;
    const float over=(float)z[1], redumax=(float)z[2];
    const float fx=(float)z[3], fy=(float)z[4], fz=(float)z[5];
    for(uint32_t p=npb;p<np;p++){
      if((code[p]&CODE_MASKSPECIAL)!=0 || CODE_TYPE(code[p])!=CODE_TYPE_FLUID)continue;
      const double x=pos[3*(size_t)p], y=pos[3*(size_t)p+1], zz=pos[3*(size_t)p+2];
      double fdis;
      if(z[0]==0){
        const float dist=(float)z[10];
        const double vdis=z[6]*x+z[7]*y+z[8]*zz+z[9];
        if(!(0<vdis && vdis<=(double)(dist+over)))continue;
        if(z[11]!=0){
          int in=(zz>=z[12] && zz<=z[13]);
          for(int q=0;q<4 && in;q++){ const double* pl=z+14+4*q; in=(pl[0]*x+pl[1]*y+pl[2]*zz+pl[3]<=0); }
          if(!in)continue;
        }
        fdis=(vdis>=(double)dist? 1.: vdis/(double)dist);
      }
      else{
        const double *m1=z+6,*m2=z+9,*x1=z+12,*x2=z+15,*o1=z+18,*o2=z+21,*b1=z+24,*b2=z+27;
        if(!(o1[0]<=x && o1[1]<=y && o1[2]<=zz && x<=o2[0] && y<=o2[1] && zz<=o2[2]))continue;
        if(m1[0]<=x && m1[1]<=y && m1[2]<=zz && x<=m2[0] && y<=m2[1] && zz<=m2[2])continue;
        fdis=1.;
        if(x1[0]<=x && x1[1]<=y && x1[2]<=zz && x<=x2[0] && y<=x2[1] && zz<=x2[2]){
          double f; fdis=0;
          if(b2[2]){ f=(zz-m2[2])/b2[2]; if(f>fdis)fdis=f; }
          if(b2[1]){ f=(y-m2[1])/b2[1]; if(f>fdis)fdis=f; }
          if(b2[0]){ f=(x-m2[0])/b2[0]; if(f>fdis)fdis=f; }
          if(b1[2]){ f=(m1[2]-zz)/b1[2]; if(f>fdis)fdis=f; }
          if(b1[1]){ f=(m1[1]-y)/b1[1]; if(f>fdis)fdis=f; }
          if(b1[0]){ f=(m1[0]-x)/b1[0]; if(f>fdis)fdis=f; }
        }
      }
      const double redudt=dt*(fdis*fdis)*redumax;
      double rx=1.-redudt*fx, ry=1.-redudt*fy, rz=1.-redudt*fz;
      if(rx<0)rx=0;
      if(ry<0)ry=0;
      if(rz<0)rz=0;
      float* v=velrho+4*(size_t)p;
      v[0]=(float)(rx*v[0]); v[1]=(float)(ry*v[1]); v[2]=(float)(rz*v[2]);
    }
  }
}

int orc_num_threads(void){
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
size_t orc_sizeof_cte(void){ return sizeof(orc_cte); }
size_t orc_sizeof_div(void){ return sizeof(orc_div); }

/* ------------------------------------------------------------------------------------------------ */
/* fsph::GetKernel_WabFac (FunSphKernel.h:71-85 cubic, :235-245 Wendland) */
static inline float kernel_wabfac(const orc_cte* c, float rr2, float* fac){
  const float rad=sqrtf(rr2), qq=rad/c->kernelh;
  if(c->tkernel==2){
    const float wqq1=1.f-0.5f*qq, wqq2=wqq1*wqq1;
    *fac=c->bwenh*wqq2*wqq1;
    const float wqq=qq+qq+1.f;
    return c->awen*wqq*wqq2*wqq2;
  }
  if(rad>c->kernelh){ const float wqq1=2.0f-qq, wqq2=wqq1*wqq1; *fac=(c->cub_c2*wqq2/rad); return c->cub_a24*(wqq2*wqq1); }
  { const float wqq2=qq*qq; *fac=((c->cub_c1*qq+c->cub_d1*wqq2)/rad); return c->cub_a2*(1.0f+(0.75f*qq-1.5f)*wqq2); }
}

/* JSphCpu::InteractionMdbcCorrectionT2<tker,sim2d=false> (JSphCpu_mdbc.cpp:36-150), SlipMode=1 (vel0).
 * boundnor holds the doubled normals (boundary -> ghost node, JSph.cpp:1417), 3 floats per particle. */
void orc_mdbc_vel0(const orc_cte* c, const orc_div* d, uint32_t n, const uint32_t* begincell, const double* pos,
  const typecode* code, const float* boundnor, float* velrho)
{
  const uint32_t* beg=begincell+d->box_fluid;
  #pragma omp parallel for schedule(guided)
  for(int64_t p1=0;p1<(int64_t)n;p1++){
    const float* nr=boundnor+3*p1;
    if(nr[0]==0 && nr[1]==0 && nr[2]==0)continue;
    float rhofinal=FLT_MAX;
    double gx=pos[3*p1]+(double)nr[0], gy=pos[3*p1+1]+(double)nr[1], gz=pos[3*p1+2]+(double)nr[2];
    if(c->periactive)periodic_wrap(c,&gx,&gy,&gz);   /* UpdatePeriodicPos(gposp1), JSphCpu_mdbc.cpp:50 */
    float rhop1=0, grx=0, gry=0, grz=0;
    double a[4][4]; memset(a,0,sizeof(a));
    /* nsearch::Init(pos,...) (JCellSearch_inline.h:54-69) */
    const int cx=(int)((gx-c->domposmin[0])/c->scell)-(int)d->cellmin[0];
    const int cy=(int)((gy-c->domposmin[1])/c->scell)-(int)d->cellmin[1];
    const int cz=(int)((gz-c->domposmin[2])/c->scell)-(int)d->cellmin[2];
    const int s=c->scelldiv, ncx=(int)d->ncx, ncy=(int)d->ncy, ncz=(int)d->ncz;
    const int xini=cx-(cx<s? cx: s), xfin=cx+(ncx-cx-1<s? ncx-cx-1: s)+1;
    const int yini=cy-(cy<s? cy: s), yfin=cy+(ncy-cy-1<s? ncy-cy-1: s)+1;
    const int zini=cz-(cz<s? cz: s), zfin=cz+(ncz-cz-1<s? ncz-cz-1: s)+1;
    if(xfin>xini && xini>=0 && yini>=0 && zini>=0)
    for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++){
      const uint32_t pini=beg[xini+y*ncx+z*(int)d->nsheet], pfin=beg[xfin+y*ncx+z*(int)d->nsheet];
      for(uint32_t p2=pini;p2<pfin;p2++){
        const float drx=(float)(gx-pos[3*(size_t)p2]), dry=(float)(gy-pos[3*(size_t)p2+1]), drz=(float)(gz-pos[3*(size_t)p2+2]);
        const float rr2=(drx*drx+dry*dry+drz*drz);
        if(rr2<=c->kernelsize2 && CODE_TYPE(code[p2])==CODE_TYPE_FLUID){
          float fac; const float wab=kernel_wabfac(c,rr2,&fac);
          const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
          const float massp2=c->massfluid, volp2=massp2/velrho[4*(size_t)p2+3];
          rhop1+=massp2*wab; grx+=massp2*frx; gry+=massp2*fry; grz+=massp2*frz;
          const float vwab=wab*volp2,vfrx=frx*volp2,vfry=fry*volp2,vfrz=frz*volp2;
          a[0][0]+=vwab; a[0][1]+=drx*vwab; a[0][2]+=dry*vwab; a[0][3]+=drz*vwab;
          a[1][0]+=vfrx; a[1][1]+=drx*vfrx; a[1][2]+=dry*vfrx; a[1][3]+=drz*vfrx;
          a[2][0]+=vfry; a[2][1]+=drx*vfry; a[2][2]+=dry*vfry; a[2][3]+=drz*vfry;
          a[3][0]+=vfrz; a[3][1]+=drx*vfrz; a[3][2]+=dry*vfrz; a[3][3]+=drz*vfrz;
        }
      }
    }
    {
      #define A(i,j) a[i-1][j-1]
      const double det=(A(1,4)*A(2,3)*A(3,2)*A(4,1) - A(1,3)*A(2,4)*A(3,2)*A(4,1) - A(1,4)*A(2,2)*A(3,3)*A(4,1) + A(1,2)*A(2,4)*A(3,3)*A(4,1) +
                        A(1,3)*A(2,2)*A(3,4)*A(4,1) - A(1,2)*A(2,3)*A(3,4)*A(4,1) - A(1,4)*A(2,3)*A(3,1)*A(4,2) + A(1,3)*A(2,4)*A(3,1)*A(4,2) +
                        A(1,4)*A(2,1)*A(3,3)*A(4,2) - A(1,1)*A(2,4)*A(3,3)*A(4,2) - A(1,3)*A(2,1)*A(3,4)*A(4,2) + A(1,1)*A(2,3)*A(3,4)*A(4,2) +
                        A(1,4)*A(2,2)*A(3,1)*A(4,3) - A(1,2)*A(2,4)*A(3,1)*A(4,3) - A(1,4)*A(2,1)*A(3,2)*A(4,3) + A(1,1)*A(2,4)*A(3,2)*A(4,3) +
                        A(1,2)*A(2,1)*A(3,4)*A(4,3) - A(1,1)*A(2,2)*A(3,4)*A(4,3) - A(1,3)*A(2,2)*A(3,1)*A(4,4) + A(1,2)*A(2,3)*A(3,1)*A(4,4) +
                        A(1,3)*A(2,1)*A(3,2)*A(4,4) - A(1,1)*A(2,3)*A(3,2)*A(4,4) - A(1,2)*A(2,1)*A(3,3)*A(4,4) + A(1,1)*A(2,2)*A(3,3)*A(4,4));
      if(fabs(det)>=(double)1e-3f){
        const double i11=(A(2,2)*(A(3,3)*A(4,4)-A(3,4)*A(4,3))+A(2,3)*(A(3,4)*A(4,2)-A(3,2)*A(4,4))+A(2,4)*(A(3,2)*A(4,3)-A(3,3)*A(4,2)))/det;
        const double i21=(A(2,1)*(A(3,4)*A(4,3)-A(3,3)*A(4,4))+A(2,3)*(A(3,1)*A(4,4)-A(3,4)*A(4,1))+A(2,4)*(A(3,3)*A(4,1)-A(3,1)*A(4,3)))/det;
        const double i31=(A(2,1)*(A(3,2)*A(4,4)-A(3,4)*A(4,2))+A(2,2)*(A(3,4)*A(4,1)-A(3,1)*A(4,4))+A(2,4)*(A(3,1)*A(4,2)-A(3,2)*A(4,1)))/det;
        const double i41=(A(2,1)*(A(3,3)*A(4,2)-A(3,2)*A(4,3))+A(2,2)*(A(3,1)*A(4,3)-A(3,3)*A(4,1))+A(2,3)*(A(3,2)*A(4,1)-A(3,1)*A(4,2)))/det;
        const double i12=(A(1,2)*(A(3,4)*A(4,3)-A(3,3)*A(4,4))+A(1,3)*(A(3,2)*A(4,4)-A(3,4)*A(4,2))+A(1,4)*(A(3,3)*A(4,2)-A(3,2)*A(4,3)))/det;
        const double i22=(A(1,1)*(A(3,3)*A(4,4)-A(3,4)*A(4,3))+A(1,3)*(A(3,4)*A(4,1)-A(3,1)*A(4,4))+A(1,4)*(A(3,1)*A(4,3)-A(3,3)*A(4,1)))/det;
        const double i32=(A(1,1)*(A(3,4)*A(4,2)-A(3,2)*A(4,4))+A(1,2)*(A(3,1)*A(4,4)-A(3,4)*A(4,1))+A(1,4)*(A(3,2)*A(4,1)-A(3,1)*A(4,2)))/det;
        const double i42=(A(1,1)*(A(3,2)*A(4,3)-A(3,3)*A(4,2))+A(1,2)*(A(3,3)*A(4,1)-A(3,1)*A(4,3))+A(1,3)*(A(3,1)*A(4,2)-A(3,2)*A(4,1)))/det;
        const double i13=(A(1,2)*(A(2,3)*A(4,4)-A(2,4)*A(4,3))+A(1,3)*(A(2,4)*A(4,2)-A(2,2)*A(4,4))+A(1,4)*(A(2,2)*A(4,3)-A(2,3)*A(4,2)))/det;
        const double i23=(A(1,1)*(A(2,4)*A(4,3)-A(2,3)*A(4,4))+A(1,3)*(A(2,1)*A(4,4)-A(2,4)*A(4,1))+A(1,4)*(A(2,3)*A(4,1)-A(2,1)*A(4,3)))/det;
        const double i33=(A(1,1)*(A(2,2)*A(4,4)-A(2,4)*A(4,2))+A(1,2)*(A(2,4)*A(4,1)-A(2,1)*A(4,4))+A(1,4)*(A(2,1)*A(4,2)-A(2,2)*A(4,1)))/det;
        const double i43=(A(1,1)*(A(2,3)*A(4,2)-A(2,2)*A(4,3))+A(1,2)*(A(2,1)*A(4,3)-A(2,3)*A(4,1))+A(1,3)*(A(2,2)*A(4,1)-A(2,1)*A(4,2)))/det;
        const double i14=(A(1,2)*(A(2,4)*A(3,3)-A(2,3)*A(3,4))+A(1,3)*(A(2,2)*A(3,4)-A(2,4)*A(3,2))+A(1,4)*(A(2,3)*A(3,2)-A(2,2)*A(3,3)))/det;
        const double i24=(A(1,1)*(A(2,3)*A(3,4)-A(2,4)*A(3,3))+A(1,3)*(A(2,4)*A(3,1)-A(2,1)*A(3,4))+A(1,4)*(A(2,1)*A(3,3)-A(2,3)*A(3,1)))/det;
        const double i34=(A(1,1)*(A(2,4)*A(3,2)-A(2,2)*A(3,4))+A(1,2)*(A(2,1)*A(3,4)-A(2,4)*A(3,1))+A(1,4)*(A(2,2)*A(3,1)-A(2,1)*A(3,2)))/det;
        const double i44=(A(1,1)*(A(2,2)*A(3,3)-A(2,3)*A(3,2))+A(1,2)*(A(2,3)*A(3,1)-A(2,1)*A(3,3))+A(1,3)*(A(2,1)*A(3,2)-A(2,2)*A(3,1)))/det;
        const float rhoghost=(float)(i11*rhop1+i12*grx+i13*gry+i14*grz);
        const float ggx=-(float)(i21*rhop1+i22*grx+i23*gry+i24*grz);
        const float ggy=-(float)(i31*rhop1+i32*grx+i33*gry+i34*grz);
        const float ggz=-(float)(i41*rhop1+i42*grx+i43*gry+i44*grz);
        rhofinal=(rhoghost+ggx*(-nr[0])+ggy*(-nr[1])+ggz*(-nr[2]));
      }
      else if(A(1,1)>0)rhofinal=(float)(rhop1/A(1,1));
      #undef A
    }
    velrho[4*p1+3]=(rhofinal!=FLT_MAX? rhofinal: c->rhopzero);
  }
}

/* 4x4 inverse by cofactors (what fmath::Determinant4x4 / InverseMatrix4x4 compute); returns the determinant. */
static double inv4(const double m[4][4], double inv[4][4]){
  double det=0;
  for(int i=0;i<4;i++)for(int j=0;j<4;j++){
    double s3[3][3]; int r=0;
    for(int a=0;a<4;a++){ if(a==i)continue; int q=0; for(int b=0;b<4;b++){ if(b==j)continue; s3[r][q++]=m[a][b]; } r++; }
    const double minor=s3[0][0]*(s3[1][1]*s3[2][2]-s3[1][2]*s3[2][1])-s3[0][1]*(s3[1][0]*s3[2][2]-s3[1][2]*s3[2][0])+s3[0][2]*(s3[1][0]*s3[2][1]-s3[1][1]*s3[2][0]);
    inv[j][i]=(((i+j)&1)? -minor: minor);   /* adjugate = transposed cofactors */
  }
  for(int j=0;j<4;j++)det+=m[0][j]*inv[j][0];
  if(det)for(int i=0;i<4;i++)for(int j=0;j<4;j++)inv[i][j]/=det;
  return det;
}

/* Interaction_Mdbc2Correction, 3-D (JSphCpu_mdbc.cpp:246-436): SlipMode 2 (no-slip) / 3 (free-slip).  For every boundary particle
 * with a normal: submerged test of the ghost node, density by the corrected kernel (matrix inverse if well conditioned, Shepard
 * otherwise), pressure cloning to the particle (Mdbc2PressClone, :182-207), velocity of the ghost node, boundmode and tangential
 * velocity.  motionace may be NULL (fixed boundaries); the motion velocity is the one the particle carries (MDBC2_KEEPVEL). */
void orc_mdbc2(const orc_cte* c, const orc_div* d, uint32_t n, const uint32_t* begincell, const double* pos, const typecode* code,
  const float* boundnor, const float* motionace, int freeslip, float* velrho, uint8_t* boundmode, float* tangenvel)
{
  const uint32_t* beg=begincell+d->box_fluid;
  #pragma omp parallel for schedule(guided)
  for(int64_t p1=0;p1<(int64_t)n;p1++){
    boundmode[p1]=0;
    const float* nr=boundnor+3*p1;
    if(nr[0]==0 && nr[1]==0 && nr[2]==0)continue;
    double gx=pos[3*p1]+(double)nr[0], gy=pos[3*p1+1]+(double)nr[1], gz=pos[3*p1+2]+(double)nr[2];
    if(c->periactive)periodic_wrap(c,&gx,&gy,&gz);
    float rhop1=0, grx=0, gry=0, grz=0, sumwab=0, submerged=0;
    double vx=0,vy=0,vz=0, a[4][4]; memset(a,0,sizeof(a));
    const int cx=(int)((gx-c->domposmin[0])/c->scell)-(int)d->cellmin[0];
    const int cy=(int)((gy-c->domposmin[1])/c->scell)-(int)d->cellmin[1];
    const int cz=(int)((gz-c->domposmin[2])/c->scell)-(int)d->cellmin[2];
    const int s=c->scelldiv, ncx=(int)d->ncx, ncy=(int)d->ncy, ncz=(int)d->ncz;
    const int xini=cx-(cx<s? cx: s), xfin=cx+(ncx-cx-1<s? ncx-cx-1: s)+1;
    const int yini=cy-(cy<s? cy: s), yfin=cy+(ncy-cy-1<s? ncy-cy-1: s)+1;
    const int zini=cz-(cz<s? cz: s), zfin=cz+(ncz-cz-1<s? ncz-cz-1: s)+1;
    if(xfin>xini && xini>=0 && yini>=0 && zini>=0)
    for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++){
      const uint32_t pini=beg[xini+y*ncx+z*(int)d->nsheet], pfin=beg[xfin+y*ncx+z*(int)d->nsheet];
      for(uint32_t p2=pini;p2<pfin;p2++){
        const float drx=(float)(gx-pos[3*(size_t)p2]), dry=(float)(gy-pos[3*(size_t)p2+1]), drz=(float)(gz-pos[3*(size_t)p2+2]);
        const float rr2=(drx*drx+dry*dry+drz*drz);
        if(rr2<=c->kernelsize2 && CODE_TYPE(code[p2])==CODE_TYPE_FLUID){
          float fac; const float wab=kernel_wabfac(c,rr2,&fac);
          const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
          const float* v2=velrho+4*(size_t)p2;
          const float massp2=c->massfluid, volp2=massp2/v2[3];
          submerged-=volp2*(drx*frx+dry*fry+drz*frz);
          rhop1+=massp2*wab; grx+=massp2*frx; gry+=massp2*fry; grz+=massp2*frz;
          const float vwab=wab*volp2,vfrx=frx*volp2,vfry=fry*volp2,vfrz=frz*volp2;
          sumwab+=vwab;
          vx+=vwab*v2[0]; vy+=vwab*v2[1]; vz+=vwab*v2[2];
          a[0][0]+=vwab; a[0][1]-=drx*vwab; a[0][2]-=dry*vwab; a[0][3]-=drz*vwab;
          a[1][0]+=vfrx; a[1][1]-=drx*vfrx; a[1][2]-=dry*vfrx; a[1][3]-=drz*vfrx;
          a[2][0]+=vfry; a[2][1]-=drx*vfry; a[2][2]-=dry*vfry; a[2][3]-=drz*vfry;
          a[3][0]+=vfrz; a[3][1]-=drx*vfrz; a[3][2]-=dry*vfrz; a[3][3]-=drz*vfrz;
        }
      }
    }
    const float snormal=sqrtf(nr[0]*nr[0]+nr[1]*nr[1]+nr[2]*nr[2]);
    const float nx=nr[0]/snormal, ny=nr[1]/snormal, nz=nr[2]/snormal;
    const float* vm=velrho+4*p1;
    float rhofinal, v2x,v2y,v2z;
    if(submerged>0.f){
      boundmode[p1]=1;
      if(sumwab<0.1f){ rhofinal=(float)(rhop1/a[0][0]); if(rhofinal<c->rhopzero)rhofinal=c->rhopzero; }
      else{
        double inv[4][4];
        const double det=inv4(a,inv);
        if(fabs(det)>=0.001){
          float na=0, ni=0;
          for(int i=0;i<4;i++){
            const float ra=(float)(fabs(a[i][0])+fabs(a[i][1])+fabs(a[i][2])+fabs(a[i][3]));
            const float ri=(float)(fabs(inv[i][0])+fabs(inv[i][1])+fabs(inv[i][2])+fabs(inv[i][3]));
            if(ra>na)na=ra;
            if(ri>ni)ni=ri;
          }
          const float condinf=(float)((double)c->dp*(double)c->dp)*na*ni;
          if(condinf<=50)rhofinal=(float)(inv[0][0]*rhop1+inv[0][1]*grx+inv[0][2]*gry+inv[0][3]*grz);
          else rhofinal=(float)(rhop1/a[0][0]);
        }
        else rhofinal=(float)(rhop1/a[0][0]);
      }
      const float pghost=(float)(c->cs0*c->cs0*(rhofinal-c->rhopzero));
      const float normpos=(-nr[0])*nx+(-nr[1])*ny+(-nr[2])*nz;
      const float ma[3]={motionace? motionace[3*p1]: 0.f, motionace? motionace[3*p1+1]: 0.f, motionace? motionace[3*p1+2]: 0.f};
      const float normforce=c->rhopzero*((c->gravity[0]-ma[0])*nx+(c->gravity[1]-ma[1])*ny+(c->gravity[2]-ma[2])*nz);
      const float pressfinal=pghost+normforce*normpos;
      rhofinal=c->rhopzero+(float)(pressfinal/(c->cs0*c->cs0));
      const float gvx=(float)(vx/a[0][0]), gvy=(float)(vy/a[0][0]), gvz=(float)(vz/a[0][0]);
      if(freeslip){ v2x=gvx; v2y=gvy; v2z=gvz; }
      else{ v2x=vm[0]+vm[0]-gvx; v2y=vm[1]+vm[1]-gvy; v2z=vm[2]+vm[2]-gvz; }
    }
    else{ boundmode[p1]=2; rhofinal=c->rhopzero; v2x=vm[0]; v2y=vm[1]; v2z=vm[2]; }
    velrho[4*p1+3]=rhofinal;     /* MDBC2_KEEPVEL: only the density is written */
    const float vdn=v2x*nx+v2y*ny+v2z*nz;
    tangenvel[3*p1]=v2x-vdn*nx; tangenvel[3*p1+1]=v2y-vdn*ny; tangenvel[3*p1+2]=v2z-vdn*nz;
  }
}

