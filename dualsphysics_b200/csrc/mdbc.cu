// mdbc.cu -- modified Dynamic Boundary Condition, ghost-node density extrapolation (SlipMode=1, "vel0").
//
// Replaces cusph::Interaction_MdbcCorrection -> KerInteractionMdbcCorrection_Fast (src/source/JSphGpu_mdbc_iker.cu:33-213),
// called from JSphGpuSingle::MdbcBoundCorrection (src/source/JSphGpuSingle.cpp:686-712).  The arithmetic follows the CPU
// statement JSphCpu::InteractionMdbcCorrectionT2 (src/source/JSphCpu_mdbc.cpp:36-150): FP32 sums of rho, grad(rho); the
// 4x4 correction matrix accumulated and inverted in FP64 (fmath::Determinant4x4 / InverseMatrix4x4, FunctionsMath.h:210-306);
// first-order result if |det|>=1e-3, zeroth order if a11>0, else rho0.  The boundary normals arrive already doubled
// (boundary -> ghost node, JSph.cpp:1417).  One thread per boundary particle; the pressure cached for the interaction
// (posp.w) is refreshed for the corrected density.  mDBC2 (SlipMode 2/3): k_mdbc2 below.
#include "b2s_internal.h"
#include <float.h>

#define M_BS 128

__device__ __forceinline__ float m_rcp(float x){ float r; asm("rcp.approx.ftz.f32 %0, %1;":"=f"(r):"f"(x)); return r; }


__global__ void __launch_bounds__(M_BS) k_mdbc_vel0(const DivState* __restrict__ ds, const uint32_t* __restrict__ begt,
  const double2* __restrict__ posxy, const double* __restrict__ posz, const typecode* __restrict__ code, const float4* __restrict__ boundnor,
  const uint32_t* __restrict__ dcell, float4* posp, float4* velrho, const __grid_constant__ DevCte c)
{
  const uint32_t p1=blockIdx.x*blockDim.x+threadIdx.x;
  const uint32_t n=(c.use_floating? ds->np: ds->npbok);    // UseNormalsFt is approximated by use_floating (normals of floatings are zero unless given)
  if(p1>=n)return;
  const float4 nor=boundnor[p1];
  if(nor.x==0.f && nor.y==0.f && nor.z==0.f)return;
  float rhofinal=FLT_MAX;
  //-Ghost node position (double, like the reference) and its cell.
  const double2 xy=posxy[p1];
  double gx=xy.x+double(nor.x), gy=xy.y+double(nor.y), gz=posz[p1]+double(nor.z);
  if(c.periactive){ uint32_t gcell; b2s_pos_to_cell(c,gx,gy,gz,gcell); }   // UpdatePeriodicPos of the ghost node (JSphCpu_mdbc.cpp:50,267)
  const int gcx=int((gx-c.domposmin[0])/c.scell_d), gcy=int((gy-c.domposmin[1])/c.scell_d), gcz=int((gz-c.domposmin[2])/c.scell_d);
  const int cx=gcx-int(ds->cellmin[0]), cy=gcy-int(ds->cellmin[1]), cz=gcz-int(ds->cellmin[2]);
  const int d=c.scelldiv, ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const int xini=cx-(cx<d? cx: d), xfin=cx+(ncx-cx-1<d? ncx-cx-1: d)+1;
  const int yini=cy-(cy<d? cy: d), yfin=cy+(ncy-cy-1<d? ncy-cy-1: d)+1;
  const int zini=cz-(cz<d? cz: d), zfin=cz+(ncz-cz-1<d? ncz-cz-1: d)+1;
  //-Ghost position relative to its own cell origin (same frame as posp).
  const float grx=float((gx-c.domposmin[0])-c.scell_d*double((gcx>>B2S_XBLK)<<B2S_XBLK));
  const float gry=float((gy-c.domposmin[1])-c.scell_d*double(gcy));
  const float grz=float((gz-c.domposmin[2])-c.scell_d*double(gcz));
  float rhop1=0.f, grdx=0.f, grdy=0.f, grdz=0.f;
  double a11=0,a12=0,a13=0,a14=0,a21=0,a22=0,a23=0,a24=0,a31=0,a32=0,a33=0,a34=0,a41=0,a42=0,a43=0,a44=0;
  const uint32_t* beg=begt+ds->box_fluid;
  if(xfin>xini && xini>=0 && yini>=0 && zini>=0)
  for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++){
    const int row=y*ncx+z*nsheet;
    const float offy=c.scell*float(cy-y), offz=c.scell*float(cz-z);
    for(int x=xini;x<xfin;x++){
      const uint32_t pini=beg[row+x], pfin=beg[row+x+1];
      const float offx=b2s_offx(c,gcx,x+int(ds->cellmin[0]));
      for(uint32_t p2=pini;p2<pfin;p2++){
        const float4 pp2=posp[p2];
        const float drx=(grx+offx)-pp2.x, dry=(gry+offy)-pp2.y, drz=(grz+offz)-pp2.z;
        const float rr2=drx*drx+dry*dry+drz*drz;
        if(rr2<=c.kernelsize2 && (!c.use_floating || B2S_CODE_ISFLUID(code[p2]))){
          float fac;
          const float wab=wab_fac(c.tkernel,rr2,c,fac);
          const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
          const float massp2=c.massf;
          const float volp2=massp2/velrho[p2].w;
          rhop1+=massp2*wab; grdx+=massp2*frx; grdy+=massp2*fry; grdz+=massp2*frz;
          const float vwab=wab*volp2, vfrx=frx*volp2, vfry=fry*volp2, vfrz=frz*volp2;
          a11+=vwab;  a12+=drx*vwab;  a13+=dry*vwab;  a14+=drz*vwab;
          a21+=vfrx;  a22+=drx*vfrx;  a23+=dry*vfrx;  a24+=drz*vfrx;
          a31+=vfry;  a32+=drx*vfry;  a33+=dry*vfry;  a34+=drz*vfry;
          a41+=vfrz;  a42+=drx*vfrz;  a43+=dry*vfrz;  a44+=drz*vfrz;
        }
      }
    }
  }
  {
    const float dpx=-nor.x, dpy=-nor.y, dpz=-nor.z;   // boundary particle - ghost node
    const double det=(a14*a23*a32*a41 - a13*a24*a32*a41 - a14*a22*a33*a41 + a12*a24*a33*a41 +
                      a13*a22*a34*a41 - a12*a23*a34*a41 - a14*a23*a31*a42 + a13*a24*a31*a42 +
                      a14*a21*a33*a42 - a11*a24*a33*a42 - a13*a21*a34*a42 + a11*a23*a34*a42 +
                      a14*a22*a31*a43 - a12*a24*a31*a43 - a14*a21*a32*a43 + a11*a24*a32*a43 +
                      a12*a21*a34*a43 - a11*a22*a34*a43 - a13*a22*a31*a44 + a12*a23*a31*a44 +
                      a13*a21*a32*a44 - a11*a23*a32*a44 - a12*a21*a33*a44 + a11*a22*a33*a44);
    if(fabs(det)>=double(1e-3f)){
      const double i11=(a22*(a33*a44-a34*a43)+a23*(a34*a42-a32*a44)+a24*(a32*a43-a33*a42))/det;
      const double i21=(a21*(a34*a43-a33*a44)+a23*(a31*a44-a34*a41)+a24*(a33*a41-a31*a43))/det;
      const double i31=(a21*(a32*a44-a34*a42)+a22*(a34*a41-a31*a44)+a24*(a31*a42-a32*a41))/det;
      const double i41=(a21*(a33*a42-a32*a43)+a22*(a31*a43-a33*a41)+a23*(a32*a41-a31*a42))/det;
      const double i12=(a12*(a34*a43-a33*a44)+a13*(a32*a44-a34*a42)+a14*(a33*a42-a32*a43))/det;
      const double i22=(a11*(a33*a44-a34*a43)+a13*(a34*a41-a31*a44)+a14*(a31*a43-a33*a41))/det;
      const double i32=(a11*(a34*a42-a32*a44)+a12*(a31*a44-a34*a41)+a14*(a32*a41-a31*a42))/det;
      const double i42=(a11*(a32*a43-a33*a42)+a12*(a33*a41-a31*a43)+a13*(a31*a42-a32*a41))/det;
      const double i13=(a12*(a23*a44-a24*a43)+a13*(a24*a42-a22*a44)+a14*(a22*a43-a23*a42))/det;
      const double i23=(a11*(a24*a43-a23*a44)+a13*(a21*a44-a24*a41)+a14*(a23*a41-a21*a43))/det;
      const double i33=(a11*(a22*a44-a24*a42)+a12*(a24*a41-a21*a44)+a14*(a21*a42-a22*a41))/det;
      const double i43=(a11*(a23*a42-a22*a43)+a12*(a21*a43-a23*a41)+a13*(a22*a41-a21*a42))/det;
      const double i14=(a12*(a24*a33-a23*a34)+a13*(a22*a34-a24*a32)+a14*(a23*a32-a22*a33))/det;
      const double i24=(a11*(a23*a34-a24*a33)+a13*(a24*a31-a21*a34)+a14*(a21*a33-a23*a31))/det;
      const double i34=(a11*(a24*a32-a22*a34)+a12*(a21*a34-a24*a31)+a14*(a22*a31-a21*a32))/det;
      const double i44=(a11*(a22*a33-a23*a32)+a12*(a23*a31-a21*a33)+a13*(a21*a32-a22*a31))/det;
      const float rhoghost=float(i11*rhop1+i12*grdx+i13*grdy+i14*grdz);
      const float ggx=-float(i21*rhop1+i22*grdx+i23*grdy+i24*grdz);
      const float ggy=-float(i31*rhop1+i32*grdx+i33*grdy+i34*grdz);
      const float ggz=-float(i41*rhop1+i42*grdx+i43*grdy+i44*grdz);
      rhofinal=(rhoghost+ggx*dpx+ggy*dpy+ggz*dpz);
    }
    else if(a11>0)rhofinal=float(double(rhop1)/a11);
  }
  rhofinal=(rhofinal!=FLT_MAX? rhofinal: c.rhopzero);
  float4 vr=velrho[p1]; vr.w=rhofinal; velrho[p1]=vr;
  float4 pp=posp[p1]; pp.w=b2s_press(rhofinal,c); posp[p1]=pp;
}

// ------------------------------------------------------------------------------------------------
// mDBC2 (SlipMode 2 no-slip / 3 free-slip): JSphCpu::InteractionMdbc2CorrectionT2 (src/source/JSphCpu_mdbc.cpp:246-436; GPU
// KerInteractionMdbc2Correction_Fast, JSphGpu_mdbc_iker.cu:289-582) with MDBC2_KEEPVEL defined as in the reference
// (DualSphDef.h:471): the boundary particle keeps its velocity, only its density is replaced.
//   submerged ghost node : boundmode=MDBC2; density from the first-order matrix if |det|>=1e-3 and the scaled condition number
//                          dp^2*|A|inf*|A^-1|inf <= 50, else Shepard (sumwab<0.1: max(rho0,Shepard)); then pressure cloning
//                          P = cs0^2 (rho_g - rho0) + rho0 ((g - a_motion).n)(dpos.n), rho = rho0 + P/cs0^2;
//                          ghost velocity by Shepard sum; tangenvel = tangential part of (2 v_motion - v_ghost) [no-slip] or v_ghost [free-slip]
//   not submerged        : boundmode=MDBC2OFF (the interaction then uses mass 0), rho=rho0, tangenvel from the motion velocity.
// Fixed boundaries only in this build: v_motion = a_motion = 0 (moving/floating + mDBC2 is rejected by the loader).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(M_BS) k_mdbc2(const DivState* __restrict__ ds, const uint32_t* __restrict__ begt,
  const double2* __restrict__ posxy, const double* __restrict__ posz, const typecode* __restrict__ code, const float4* __restrict__ boundnor,
  float4* posp, float4* velrho, uint8_t* boundmode, float4* tangenvel, const float4* __restrict__ motionace, int freeslip, const __grid_constant__ DevCte c)
{
  const uint32_t p1=blockIdx.x*blockDim.x+threadIdx.x;
  if(p1>=ds->npb)return;
  boundmode[p1]=B2S_BMODE_DBC;
  const float4 nor=boundnor[p1];
  if(nor.x==0.f && nor.y==0.f && nor.z==0.f)return;
  const double2 xy=posxy[p1];
  double gx=xy.x+double(nor.x), gy=xy.y+double(nor.y), gz=posz[p1]+double(nor.z);
  if(c.periactive){ uint32_t gcell; b2s_pos_to_cell(c,gx,gy,gz,gcell); }   // UpdatePeriodicPos of the ghost node (JSphCpu_mdbc.cpp:50,267)
  const int gcx=int((gx-c.domposmin[0])/c.scell_d), gcy=int((gy-c.domposmin[1])/c.scell_d), gcz=int((gz-c.domposmin[2])/c.scell_d);
  const int cx=gcx-int(ds->cellmin[0]), cy=gcy-int(ds->cellmin[1]), cz=gcz-int(ds->cellmin[2]);
  const int d=c.scelldiv, ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const int xini=cx-(cx<d? cx: d), xfin=cx+(ncx-cx-1<d? ncx-cx-1: d)+1;
  const int yini=cy-(cy<d? cy: d), yfin=cy+(ncy-cy-1<d? ncy-cy-1: d)+1;
  const int zini=cz-(cz<d? cz: d), zfin=cz+(ncz-cz-1<d? ncz-cz-1: d)+1;
  const float grx=float((gx-c.domposmin[0])-c.scell_d*double((gcx>>B2S_XBLK)<<B2S_XBLK));
  const float gry=float((gy-c.domposmin[1])-c.scell_d*double(gcy));
  const float grz=float((gz-c.domposmin[2])-c.scell_d*double(gcz));
  float rhop1=0.f, grdx=0.f, grdy=0.f, grdz=0.f, sumwab=0.f, submerged=0.f;
  double vx=0,vy=0,vz=0;
  double a11=0,a12=0,a13=0,a14=0,a21=0,a22=0,a23=0,a24=0,a31=0,a32=0,a33=0,a34=0,a41=0,a42=0,a43=0,a44=0;
  const uint32_t* beg=begt+ds->box_fluid;
  if(xfin>xini && xini>=0 && yini>=0 && zini>=0)
  for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++){
    const int row=y*ncx+z*nsheet;
    const float offy=c.scell*float(cy-y), offz=c.scell*float(cz-z);
    for(int x=xini;x<xfin;x++){
      const uint32_t pini=beg[row+x], pfin=beg[row+x+1];
      const float offx=b2s_offx(c,gcx,x+int(ds->cellmin[0]));
      for(uint32_t p2=pini;p2<pfin;p2++){
        const float4 pp2=posp[p2];
        const float drx=(grx+offx)-pp2.x, dry=(gry+offy)-pp2.y, drz=(grz+offz)-pp2.z;
        const float rr2=drx*drx+dry*dry+drz*drz;
        if(rr2<=c.kernelsize2 && B2S_CODE_ISFLUID(code[p2])){
          float fac;
          const float wab=wab_fac(c.tkernel,rr2,c,fac);
          const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
          const float4 v2=velrho[p2];
          const float massp2=c.massf;
          const float volp2=massp2/v2.w;
          submerged-=volp2*(drx*frx+dry*fry+drz*frz);
          rhop1+=massp2*wab; grdx+=massp2*frx; grdy+=massp2*fry; grdz+=massp2*frz;
          const float vwab=wab*volp2, vfrx=frx*volp2, vfry=fry*volp2, vfrz=frz*volp2;
          sumwab+=vwab;
          vx+=vwab*v2.x; vy+=vwab*v2.y; vz+=vwab*v2.z;
          a11+=vwab;  a12-=drx*vwab;  a13-=dry*vwab;  a14-=drz*vwab;
          a21+=vfrx;  a22-=drx*vfrx;  a23-=dry*vfrx;  a24-=drz*vfrx;
          a31+=vfry;  a32-=drx*vfry;  a33-=dry*vfry;  a34-=drz*vfry;
          a41+=vfrz;  a42-=drx*vfrz;  a43-=dry*vfrz;  a44-=drz*vfrz;
        }
      }
    }
  }
  const float snormal=sqrtf(nor.x*nor.x+nor.y*nor.y+nor.z*nor.z);
  const float nx=nor.x/snormal, ny=nor.y/snormal, nz=nor.z/snormal;
  float rhofinal, tvx,tvy,tvz;
  if(submerged>0.f){
    boundmode[p1]=B2S_BMODE_MDBC2;
    if(sumwab<0.1f)rhofinal=fmaxf(c.rhopzero,float(double(rhop1)/a11));
    else{
      const double det=(a14*a23*a32*a41 - a13*a24*a32*a41 - a14*a22*a33*a41 + a12*a24*a33*a41 +
                        a13*a22*a34*a41 - a12*a23*a34*a41 - a14*a23*a31*a42 + a13*a24*a31*a42 +
                        a14*a21*a33*a42 - a11*a24*a33*a42 - a13*a21*a34*a42 + a11*a23*a34*a42 +
                        a14*a22*a31*a43 - a12*a24*a31*a43 - a14*a21*a32*a43 + a11*a24*a32*a43 +
                        a12*a21*a34*a43 - a11*a22*a34*a43 - a13*a22*a31*a44 + a12*a23*a31*a44 +
                        a13*a21*a32*a44 - a11*a23*a32*a44 - a12*a21*a33*a44 + a11*a22*a33*a44);
      if(fabs(det)>=0.001){
        const double i11=(a22*(a33*a44-a34*a43)+a23*(a34*a42-a32*a44)+a24*(a32*a43-a33*a42))/det;
        const double i21=(a21*(a34*a43-a33*a44)+a23*(a31*a44-a34*a41)+a24*(a33*a41-a31*a43))/det;
        const double i31=(a21*(a32*a44-a34*a42)+a22*(a34*a41-a31*a44)+a24*(a31*a42-a32*a41))/det;
        const double i41=(a21*(a33*a42-a32*a43)+a22*(a31*a43-a33*a41)+a23*(a32*a41-a31*a42))/det;
        const double i12=(a12*(a34*a43-a33*a44)+a13*(a32*a44-a34*a42)+a14*(a33*a42-a32*a43))/det;
        const double i22=(a11*(a33*a44-a34*a43)+a13*(a34*a41-a31*a44)+a14*(a31*a43-a33*a41))/det;
        const double i32=(a11*(a34*a42-a32*a44)+a12*(a31*a44-a34*a41)+a14*(a32*a41-a31*a42))/det;
        const double i42=(a11*(a32*a43-a33*a42)+a12*(a33*a41-a31*a43)+a13*(a31*a42-a32*a41))/det;
        const double i13=(a12*(a23*a44-a24*a43)+a13*(a24*a42-a22*a44)+a14*(a22*a43-a23*a42))/det;
        const double i23=(a11*(a24*a43-a23*a44)+a13*(a21*a44-a24*a41)+a14*(a23*a41-a21*a43))/det;
        const double i33=(a11*(a22*a44-a24*a42)+a12*(a24*a41-a21*a44)+a14*(a21*a42-a22*a41))/det;
        const double i43=(a11*(a23*a42-a22*a43)+a12*(a21*a43-a23*a41)+a13*(a22*a41-a21*a42))/det;
        const double i14=(a12*(a24*a33-a23*a34)+a13*(a22*a34-a24*a32)+a14*(a23*a32-a22*a33))/det;
        const double i24=(a11*(a23*a34-a24*a33)+a13*(a24*a31-a21*a34)+a14*(a21*a33-a23*a31))/det;
        const double i34=(a11*(a24*a32-a22*a34)+a12*(a21*a34-a24*a31)+a14*(a22*a31-a21*a32))/det;
        const double i44=(a11*(a22*a33-a23*a32)+a12*(a23*a31-a21*a33)+a13*(a21*a32-a22*a31))/det;
        // scaled condition number with infinity norms (Mdbc2InfNorm4x4, JSphCpu_mdbc.cpp:222-229)
        const float na=fmaxf(fmaxf(float(fabs(a11)+fabs(a12)+fabs(a13)+fabs(a14)),float(fabs(a21)+fabs(a22)+fabs(a23)+fabs(a24))),
                             fmaxf(float(fabs(a31)+fabs(a32)+fabs(a33)+fabs(a34)),float(fabs(a41)+fabs(a42)+fabs(a43)+fabs(a44))));
        const float ni=fmaxf(fmaxf(float(fabs(i11)+fabs(i12)+fabs(i13)+fabs(i14)),float(fabs(i21)+fabs(i22)+fabs(i23)+fabs(i24))),
                             fmaxf(float(fabs(i31)+fabs(i32)+fabs(i33)+fabs(i34)),float(fabs(i41)+fabs(i42)+fabs(i43)+fabs(i44))));
        const float condinf=c.dp2*na*ni;
        if(condinf<=50.f)rhofinal=float(i11*rhop1+i12*grdx+i13*grdy+i14*grdz);
        else rhofinal=float(double(rhop1)/a11);
      }
      else rhofinal=float(double(rhop1)/a11);
    }
    // pressure cloning (Mdbc2PressClone, JSphCpu_mdbc.cpp:182-207); dpos = boundary particle - ghost node = -normal
    const float pghost=float(c.cs0_d*c.cs0_d*double(rhofinal-c.rhopzero));
    const float normpos=(-nor.x)*nx+(-nor.y)*ny+(-nor.z)*nz;
    // force = gravity - acceleration of the imposed motion (zero for fixed boundaries; MotionAce of the last RunMotion otherwise)
    const float4 mace=(motionace? motionace[p1]: make_float4(0.f,0.f,0.f,0.f));
    const float normforce=c.rhopzero*((c.gravity[0]-mace.x)*nx+(c.gravity[1]-mace.y)*ny+(c.gravity[2]-mace.z)*nz);
    const float pressfinal=pghost+normforce*normpos;
    rhofinal=c.rhopzero+float(double(pressfinal)/(c.cs0_d*c.cs0_d));
    const float gvx=float(vx/a11), gvy=float(vy/a11), gvz=float(vz/a11);
    // no-slip: v2 = 2*v_motion - v_ghost ; free-slip: v2 = v_ghost.  v_motion = the velocity the boundary particle carries
    // (MDBC2_KEEPVEL, DualSphDef.h:471: velrho of a boundary particle is never overwritten, so it equals MotionVel)
    const float4 vm=velrho[p1];
    const float v2x=(freeslip? gvx: vm.x+vm.x-gvx), v2y=(freeslip? gvy: vm.y+vm.y-gvy), v2z=(freeslip? gvz: vm.z+vm.z-gvz);
    const float vdn=v2x*nx+v2y*ny+v2z*nz;
    tvx=v2x-vdn*nx; tvy=v2y-vdn*ny; tvz=v2z-vdn*nz;
  }
  else{
    boundmode[p1]=B2S_BMODE_MDBC2OFF;
    rhofinal=c.rhopzero;
    const float4 vm=velrho[p1];   // tangential part of the motion velocity
    const float vdn=vm.x*nx+vm.y*ny+vm.z*nz;
    tvx=vm.x-vdn*nx; tvy=vm.y-vdn*ny; tvz=vm.z-vdn*nz;
  }
  float4 vr=velrho[p1]; vr.w=rhofinal; velrho[p1]=vr;
  float4 pp=posp[p1]; pp.w=b2s_press(rhofinal,c); posp[p1]=pp;
  tangenvel[p1]=make_float4(tvx,tvy,tvz,0.f);
}

int b2s_mdbc_impl(b2s_ctx* ctx){
  if(!ctx->boundnor)return b2s_fail(ctx,B2S_ERR_STATE,"mDBC needs boundary normals");
  if(ctx->prm.slipmode>=B2S_SLIP_NOSLIP){
    if(ctx->cte.use_floating)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"mDBC2 (SlipMode 2/3) with floating bodies is not supported");
    if(ctx->prm.casenmoving && ctx->slab_on)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"mDBC2 (SlipMode 2/3) with moving boundaries is not supported with slab decomposition");
    if(!ctx->npb_host)return B2S_OK;
    k_mdbc2<<<b2s_blocks(ctx->npb_host,M_BS),M_BS,0,ctx->stream>>>(ctx->div,(const uint32_t*)ctx->begincell,ctx->cur.posxy,ctx->cur.posz,ctx->cur.code,
      ctx->boundnor,ctx->posp,ctx->cur.velrho,ctx->boundmode,ctx->tangenvel,ctx->motionace,ctx->prm.slipmode==B2S_SLIP_FREESLIP? 1: 0,ctx->cte); B2S_LAUNCHED(ctx);
    B2S_CUDA(ctx,cudaGetLastError());
    return B2S_OK;
  }
  const uint32_t n=(ctx->cte.use_floating? ctx->np_host: ctx->npb_host);
  if(!n)return B2S_OK;
  k_mdbc_vel0<<<b2s_blocks(n,M_BS),M_BS,0,ctx->stream>>>(ctx->div,(const uint32_t*)ctx->begincell,ctx->cur.posxy,ctx->cur.posz,ctx->cur.code,
    ctx->boundnor,ctx->cur.dcell,ctx->posp,ctx->cur.velrho,ctx->cte); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
