// periodic.cu -- periodic boundaries: duplicates of the particles next to a periodic face, created before every cell divide.
//
// Replaces (under /root/reference/src/source):
//   JSphGpuSingle::RunPeriodic                    JSphGpuSingle.cpp:276-366   (CPU statement JSphCpuSingle.cpp:416-504)
//   cusph::PeriodicIgnore / PeriodicMakeList      JSphGpu_ker.cu:2105-2260    (CPU JSphCpuSingle.cpp:210-247,396-401)
//   cusph::PeriodicDuplicateVerlet/Symplectic/Normals  JSphGpu_ker.cu:2267-2399 (CPU JSphCpuSingle.cpp:258-372)
// The wrap-around of a particle that crosses a periodic face is part of the position update (b2s_pos_to_cell, b2s_internal.h).
//
// A duplicate is a full copy (id, velrho, predictor / Verlet / SPS / normal arrays) displaced by +-PeriNinc and marked
// CODE_PERIODIC: a neighbour only -- the interaction kernels skip it as a target, the integrators freeze it and mark it
// CODE_OUTIGNORE, the next divide drops it -- exactly the life cycle of the slab halo ghosts (slab.cu), whose code it shares.
//
// Order is part of the contract: the cell sort is stable, so the order in which duplicates are appended decides the particle
// order inside every cell and with it the summation order of the interaction.  The reference builds its list serially:
// boundary then fluid; per type the axes X, Y, Z; per axis first the duplicates of this type made so far (corner copies),
// then the original particles; per particle the +inc copy before the -inc copy.  Every pass here is flag -> exclusive scan ->
// scatter, which reproduces that order exactly (tables and idp order are compared bit for bit with the reference's dumps).
#include "b2s_internal.h"

#define P_BS 256

void b2s_launch_scan_excl(cudaStream_t s, uint32_t* data, uint32_t n);   // celldiv.cu

struct PeriArrays {
  uint32_t* idp; typecode* code; uint32_t* dcell; double2* posxy; double* posz; float4* velrho;
  double2* pre_xy; double* pre_z; float4* pre_v; float4* m1; float* tau; float4* nor; float4* mace;
};

// PeriodicIgnore + the bookkeeping of the head of RunPeriodic.
__global__ void __launch_bounds__(P_BS) k_peri_ignore(const DivState* __restrict__ ds, typecode* code){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p<ds->np && B2S_CODE_SPECIAL(code[p])==B2S_CODE_PERIODIC)code[p]=typecode(B2S_CODE_SETNORMAL(code[p])|B2S_CODE_OUTIGNORE);
}
__global__ void k_peri_begin(DivState* ds){ ds->peri_np0=ds->np; ds->peri_npb0=ds->npb; ds->peri_nbper=0; ds->peri_nfper=0; }

// Source range of a pass.  cblock 1 = the particles present at entry, selected by TYPE (boundary / fluid+floating) rather than by
// the index ranges [0,Npb) / [Npb,Np) of the reference: identical for a single domain, and still right in slab mode where halo
// ghosts and migrants of both types sit behind the owned particles.  cblock 0 = the duplicates of this type made so far
// (contiguous: all boundary duplicates are created before the first fluid one).
__device__ __forceinline__ void peri_range(const DivState* ds,int ctype,int cblock,uint32_t& pini,uint32_t& num){
  if(cblock){ pini=0u; num=ds->peri_np0; }
  else{ num=(ctype? ds->peri_nfper: ds->peri_nbper); pini=ds->np-num; }
}
__device__ __forceinline__ bool peri_type_ok(typecode c,int ctype){ return (B2S_CODE_TYPE(c)<B2S_CODE_TYPE_FLOATING)==(ctype==0); }

// PeriodicMakeList: bit0 = copy at +inc lies inside the map, bit1 = copy at -inc does.
__global__ void __launch_bounds__(P_BS) k_peri_flag(const DivState* __restrict__ ds,int ctype,int cblock,double ix,double iy,double iz,
  double mnx,double mny,double mnz,double mxx,double mxy,double mxz,
  const typecode* __restrict__ code,const double2* __restrict__ posxy,const double* __restrict__ posz,uint8_t* flags,uint32_t* counts)
{
  __shared__ uint32_t cnt;
  if(threadIdx.x==0)cnt=0;
  __syncthreads();
  uint32_t pini,num; peri_range(ds,ctype,cblock,pini,num);
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  uint32_t f=0;
  if(i<num){
    const uint32_t p=pini+i;
    const typecode cod=code[p];
    if(B2S_CODE_SPECIAL(cod)<=B2S_CODE_PERIODIC && peri_type_ok(cod,ctype)){
      const double2 xy=posxy[p]; const double z=posz[p];
      double x2=xy.x+ix, y2=xy.y+iy, z2=z+iz;
      if(mnx<=x2 && mny<=y2 && mnz<=z2 && x2<mxx && y2<mxy && z2<mxz)f|=1;
      x2=xy.x-ix; y2=xy.y-iy; z2=z-iz;
      if(mnx<=x2 && mny<=y2 && mnz<=z2 && x2<mxx && y2<mxy && z2<mxz)f|=2;
    }
    flags[i]=uint8_t(f);
  }
  const uint32_t n=__popc(f);
  const uint32_t w=__reduce_add_sync(0xffffffffu,n);
  if((threadIdx.x&31)==0 && w)atomicAdd(&cnt,w);
  __syncthreads();
  if(threadIdx.x==0)counts[blockIdx.x]=cnt;
}

// PeriodicDuplicatePos + PeriodicDuplicateSymplectic/Verlet/Normals for one copy.
__device__ __forceinline__ void peri_copy(const PeriArrays& a,const DevCte& c,uint32_t pnew,uint32_t pcopy,double ix,double iy,double iz){
  const double2 xy=a.posxy[pcopy]; const double z=a.posz[pcopy];
  const double px=xy.x+ix, py=xy.y+iy, pz=z+iz;
  uint32_t cx=uint32_t((px-c.domposmin[0])/c.scell_d), cy=uint32_t((py-c.domposmin[1])/c.scell_d), cz=uint32_t((pz-c.domposmin[2])/c.scell_d);
  cx=(cx<=c.domcells[0]? cx: c.domcells[0]); cy=(cy<=c.domcells[1]? cy: c.domcells[1]); cz=(cz<=c.domcells[2]? cz: c.domcells[2]);
  a.posxy[pnew]=make_double2(px,py); a.posz[pnew]=pz;
  a.dcell[pnew]=B2S_DCEL_CELL(c.domcellcode,cx,cy,cz);
  a.idp[pnew]=a.idp[pcopy];
  a.code[pnew]=typecode(B2S_CODE_SETNORMAL(a.code[pcopy])|B2S_CODE_PERIODIC);
  a.velrho[pnew]=a.velrho[pcopy];
  if(a.pre_v){ a.pre_xy[pnew]=a.pre_xy[pcopy]; a.pre_z[pnew]=a.pre_z[pcopy]; a.pre_v[pnew]=a.pre_v[pcopy]; }
  if(a.m1)a.m1[pnew]=a.m1[pcopy];
  if(a.tau){ const float2* s=reinterpret_cast<const float2*>(a.tau+size_t(pcopy)*6); float2* t=reinterpret_cast<float2*>(a.tau+size_t(pnew)*6); t[0]=s[0]; t[1]=s[1]; t[2]=s[2]; }
  if(a.nor)a.nor[pnew]=a.nor[pcopy];
  if(a.mace)a.mace[pnew]=a.mace[pcopy];
}

__global__ void __launch_bounds__(P_BS) k_peri_dup(DivState* ds,int ctype,int cblock,double ix,double iy,double iz,
  const uint8_t* __restrict__ flags,const uint32_t* __restrict__ offs,PeriArrays a,uint32_t cap,const __grid_constant__ DevCte c)
{
  __shared__ uint32_t wsum[P_BS/32];
  uint32_t pini,num; peri_range(ds,ctype,cblock,pini,num);
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  const uint32_t f=(i<num? flags[i]: 0u);
  const uint32_t n=__popc(f);
  uint32_t x=n;
  #pragma unroll
  for(int o=1;o<32;o<<=1){ const uint32_t y=__shfl_up_sync(0xffffffffu,x,o); if(lane>=o)x+=y; }
  if(lane==31)wsum[w]=x;
  __syncthreads();
  uint32_t base=offs[blockIdx.x];
  for(int k=0;k<w;k++)base+=wsum[k];
  uint32_t pnew=ds->np+base+(x-n);
  if(n && pnew+n>cap){ ds->error=2u; return; }   // capacity exceeded: reported by the host (b2s_periodic_impl)
  const uint32_t p=pini+i;
  if(f&1){ peri_copy(a,c,pnew,p, ix, iy, iz); pnew++; }
  if(f&2){ peri_copy(a,c,pnew,p,-ix,-iy,-iz); }
}
__global__ void k_peri_commit(DivState* ds,int ctype,const uint32_t* total){
  const uint32_t n=*total;
  if(ds->error==2u)return;
  ds->np+=n;
  if(ctype){ ds->peri_nfper+=n; } else{ ds->peri_nbper+=n; ds->npb+=n; }
}

int b2s_periodic_impl(b2s_ctx* ctx){
  const DevCte& c=ctx->cte;
  if(!c.periactive)return B2S_OK;
  cudaStream_t s=ctx->stream;
  const uint32_t np0=ctx->np_host, npb0=ctx->npb_host;   // exact: periodic runs refresh the host mirrors after every divide
  if(!np0)return B2S_OK;
  if(!ctx->slab_counts){
    const size_t nb=b2s_blocks(ctx->cap,P_BS);
    ctx->slab_counts_len=4*(nb+1);
    B2S_CUDA(ctx,cudaMalloc((void**)&ctx->slab_counts,ctx->slab_counts_len*4));
    B2S_CUDA(ctx,cudaMalloc((void**)&ctx->slab_flags,size_t(ctx->cap)*2));
  }
  // slab mode: the CODE_PERIODIC particles present now are the halo ghosts just received (the old duplicates were marked by the
  // integrators, or by b2s_slab_config / b2s_slab_drop_ghosts): they must survive, and they are duplicated like owned particles
  if(!ctx->slab_on){ k_peri_ignore<<<b2s_blocks(np0,P_BS),P_BS,0,s>>>(ctx->div,ctx->cur.code); B2S_LAUNCHED(ctx); }
  k_peri_begin<<<1,1,0,s>>>(ctx->div); B2S_LAUNCHED(ctx);
  PeriArrays a;
  a.idp=ctx->cur.idp; a.code=ctx->cur.code; a.dcell=ctx->cur.dcell; a.posxy=ctx->cur.posxy; a.posz=ctx->cur.posz; a.velrho=ctx->cur.velrho;
  a.pre_xy=(ctx->pre_active? ctx->pospre_xy: NULL); a.pre_z=(ctx->pre_active? ctx->pospre_z: NULL); a.pre_v=(ctx->pre_active? ctx->velrhopre: NULL);
  a.m1=ctx->velrhom1; a.tau=ctx->spstau; a.nor=ctx->boundnor; a.mace=ctx->motionace;
  const b2s_params& q=ctx->prm;
  uint32_t nper[2]={0,0}, np=np0;
  for(int ctype=0;ctype<2;ctype++)for(int ax=0;ax<3;ax++)if((c.periactive>>ax)&1){
    const double* inc=c.perinc[ax];
    for(int cblock=0;cblock<2;cblock++){
      const uint32_t num=(cblock? np0: nper[ctype]);
      if(!num)continue;
      const uint32_t nb=b2s_blocks(num,P_BS);
      k_peri_flag<<<nb,P_BS,0,s>>>(ctx->div,ctype,cblock,inc[0],inc[1],inc[2],q.map_posmin[0],q.map_posmin[1],q.map_posmin[2],
        q.map_posmax[0],q.map_posmax[1],q.map_posmax[2],ctx->cur.code,ctx->cur.posxy,ctx->cur.posz,ctx->slab_flags,ctx->slab_counts); B2S_LAUNCHED(ctx);
      B2S_CUDA(ctx,cudaMemsetAsync(ctx->slab_counts+nb,0,4,s));
      b2s_launch_scan_excl(s,ctx->slab_counts,nb+1); B2S_LAUNCHED(ctx);
      k_peri_dup<<<nb,P_BS,0,s>>>(ctx->div,ctype,cblock,inc[0],inc[1],inc[2],ctx->slab_flags,ctx->slab_counts,a,ctx->cap,ctx->cte); B2S_LAUNCHED(ctx);
      k_peri_commit<<<1,1,0,s>>>(ctx->div,ctype,ctx->slab_counts+nb); B2S_LAUNCHED(ctx);
      uint32_t cnt=0;
      B2S_CUDA(ctx,cudaMemcpyAsync(&cnt,ctx->slab_counts+nb,4,cudaMemcpyDeviceToHost,s));
      B2S_CUDA(ctx,cudaStreamSynchronize(s));
      if(uint64_t(np)+cnt>ctx->cap)return b2s_fail(ctx,B2S_ERR_NOMEM,"particle capacity %u exceeded by %u periodic duplicates (np=%u)",ctx->cap,cnt,np);
      nper[ctype]+=cnt; np+=cnt;
    }
  }
  ctx->np_host=np; ctx->npb_host=npb0+nper[0]; ctx->npf_host=ctx->np_host-ctx->npb_host;
  ctx->bound_changed=true;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
