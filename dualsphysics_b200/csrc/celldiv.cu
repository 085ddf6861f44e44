// celldiv.cu -- cell-linked neighbour list for sm_100a.
//
// Replaces JSphGpuSingle::RunCellDivide (src/source/JSphGpuSingle.cpp:372-498) and everything under it:
//   JCellDivGpuSingle::Divide/CalcCellDomain/MergeMapCellBoundFluid/PrepareNct (JCellDivGpuSingle.cpp:47-235)
//   cudiv::LimitsCell (JCellDivGpu_ker.cu:362-526)        -> k_limits (warp redux + 6 atomics)
//   cudiv::PreSortFull (JCellDivGpuSingle_ker.cu:41-109)    -> k_presort
//   cudiv::Sort = thrust::stable_sort_by_key (JCellDivGpu_ker.cu:120-127) -> hand-written stable LSD radix sort
//   cudiv::CalcBeginEndCell (JCellDivGpu_ker.cu:533-570)     -> k_beginend
//   cudiv::SortDataParticles x9 (JCellDivGpu_ker.cu:577-711) + cusphs::UpdatePosCell (JSphGpuSimple_ker.cu:41-69)
//     + cusph::ComputeVelMod/ReduMaxFloat (JSphGpu_ker.cu:295,75) + CalcRidp (:1582)      -> ONE gather kernel k_reorder
// Design: no host round trip anywhere in the divide.  The cell domain, Nct, the box layout and the
// particle counts live in a device-resident DivState that the later kernels read; the host mirrors it
// once per step.  The divide is always "full" (boundary re-sorted too); the result is identical to the
// reference's fluid-only divide because its boundary block is already in cell order (JCellDivGpuSingle.cpp:193-197).
#include "b2s_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_fp16.h>

#define DIV_BS 256

// ------------------------------------------------------------------------------------------------
// Limits: min/max cell of valid particles (code special < OUTIGNORE), LimitsCellBound/Fluid
// (JCellDivCpu.cpp:269-352 is the readable statement of the same loop).
// ------------------------------------------------------------------------------------------------
__global__ void k_limits_reset(DivState* ds, int resetbound){
  if(threadIdx.x<3){ ds->flim[threadIdx.x]=0xffffffffu; ds->flim[3+threadIdx.x]=0u;
    if(resetbound){ ds->blim[threadIdx.x]=0xffffffffu; ds->blim[3+threadIdx.x]=0u; } }
}

__global__ void __launch_bounds__(DIV_BS) k_limits(uint32_t n, uint32_t pini, const uint32_t* __restrict__ dcell,
  const typecode* __restrict__ code, uint32_t dcc, uint32_t* lim)
{
  uint32_t mn[3]={0xffffffffu,0xffffffffu,0xffffffffu}, mx[3]={0,0,0};
  #pragma unroll 8
  for(uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; i<n; i+=gridDim.x*blockDim.x){
    const uint32_t p=pini+i;
    const typecode c=code[p];
    const uint32_t rc=dcell[p];          // unconditional: the loads of the unrolled iterations go out together
    if(B2S_CODE_SPECIAL(c)<B2S_CODE_OUTIGNORE){
      const uint32_t cx=B2S_DCEL_CELLX(dcc,rc), cy=B2S_DCEL_CELLY(dcc,rc), cz=B2S_DCEL_CELLZ(dcc,rc);
      mn[0]=min(mn[0],cx); mn[1]=min(mn[1],cy); mn[2]=min(mn[2],cz);
      mx[0]=max(mx[0],cx); mx[1]=max(mx[1],cy); mx[2]=max(mx[2],cz);
    }
  }
  #pragma unroll
  for(int k=0;k<3;k++){
    const uint32_t a=__reduce_min_sync(0xffffffffu,mn[k]);
    const uint32_t b=__reduce_max_sync(0xffffffffu,mx[k]);
    if((threadIdx.x&31)==0){ if(a!=0xffffffffu)atomicMin(&lim[k],a); if(b)atomicMax(&lim[3+k],b); }
  }
}

// Slab mode: ghost/migrant particles are appended behind the owned ones, so boundary and fluid particles are no longer
// two index ranges; the limits are then selected by particle type in one pass.
__global__ void __launch_bounds__(DIV_BS) k_limits_bytype(DivState* ds, const uint32_t* __restrict__ dcell,
  const typecode* __restrict__ code, uint32_t dcc)
{
  const uint32_t n=ds->np;
  uint32_t mn[2][3]={{0xffffffffu,0xffffffffu,0xffffffffu},{0xffffffffu,0xffffffffu,0xffffffffu}}, mx[2][3]={{0,0,0},{0,0,0}};
  for(uint32_t p=blockIdx.x*blockDim.x+threadIdx.x; p<n; p+=gridDim.x*blockDim.x){
    const typecode c=code[p];
    if(B2S_CODE_SPECIAL(c)<B2S_CODE_OUTIGNORE){
      const int t=(B2S_CODE_TYPE(c)<B2S_CODE_TYPE_FLOATING? 0: 1);
      const uint32_t rc=dcell[p];
      const uint32_t cx=B2S_DCEL_CELLX(dcc,rc), cy=B2S_DCEL_CELLY(dcc,rc), cz=B2S_DCEL_CELLZ(dcc,rc);
      mn[t][0]=min(mn[t][0],cx); mn[t][1]=min(mn[t][1],cy); mn[t][2]=min(mn[t][2],cz);
      mx[t][0]=max(mx[t][0],cx); mx[t][1]=max(mx[t][1],cy); mx[t][2]=max(mx[t][2],cz);
    }
  }
  #pragma unroll
  for(int t=0;t<2;t++){
    uint32_t* lim=(t==0? ds->blim: ds->flim);
    #pragma unroll
    for(int k=0;k<3;k++){
      const uint32_t a=__reduce_min_sync(0xffffffffu,mn[t][k]);
      const uint32_t b=__reduce_max_sync(0xffffffffu,mx[t][k]);
      if((threadIdx.x&31)==0){ if(a!=0xffffffffu)atomicMin(&lim[k],a); if(b)atomicMax(&lim[3+k],b); }
    }
  }
}

// CalcCellDomain + MergeMapCellBoundFluid + PrepareNct (JCellDivGpuSingle.cpp:47-125), one thread.
__global__ void k_divstate(DivState* ds, uint32_t dcx, uint32_t dcy, uint32_t dcz, uint32_t scelldiv){
  if(threadIdx.x||blockIdx.x)return;
  const uint32_t nptot=ds->np;   // particles entering the divide = survivors of the previous one (device-side count)
  const uint32_t dom[3]={dcx,dcy,dcz};
  uint32_t cbmin[3],cbmax[3],cfmin[3],cfmax[3];
  const bool bnone=(ds->blim[0]>ds->blim[3]);   // no valid particles: min>max (cmin.x>cmax.x test)
  const bool fnone=(ds->flim[0]>ds->flim[3]);
  for(int k=0;k<3;k++){
    cbmin[k]=(bnone? dom[k]: ds->blim[k]);  cbmax[k]=(bnone? 0u: ds->blim[3+k]);
    cfmin[k]=(fnone? dom[k]: ds->flim[k]);  cfmax[k]=(fnone? 0u: ds->flim[3+k]);
  }
  uint32_t cmin[3],cmax[3];
  for(int k=0;k<3;k++){
    cmin[k]=max(min(cbmin[k],cfmin[k]),(cfmin[k]>=scelldiv? cfmin[k]-scelldiv: 0u));
    cmax[k]=min(max(cbmax[k],cfmax[k]),cfmax[k]+scelldiv);
    if(cmax[k]>=dom[k])cmax[k]=dom[k]-1;
  }
  if(cmin[0]>cmax[0] || cmin[1]>cmax[1] || cmin[2]>cmax[2]){ for(int k=0;k<3;k++)cmin[k]=cmax[k]=0; }
  for(int k=0;k<3;k++){ ds->cellmin[k]=cmin[k]; ds->cellmax[k]=cmax[k]; }
  ds->ncx=cmax[0]-cmin[0]+1; ds->ncy=cmax[1]-cmin[1]+1; ds->ncz=cmax[2]-cmin[2]+1;
  ds->nsheet=ds->ncx*ds->ncy; ds->nct=ds->nsheet*ds->ncz; ds->nctt=ds->nct*2+5+1;
  ds->box_boundignore=ds->nct;
  ds->box_fluid=ds->box_boundignore+1;
  ds->box_boundout=ds->box_fluid+ds->nct;
  ds->box_fluidout=ds->box_boundout+1;
  ds->box_boundoutignore=ds->box_fluidout+1;
  ds->box_fluidoutignore=ds->box_boundoutignore+1;
  ds->nptot=nptot;
}

// PreSortFull (JCellDivGpuSingle_ker.cu:41-109; CPU statement JCellDivCpuSingle.cpp:137-170).
__global__ void __launch_bounds__(DIV_BS) k_presort(const DivState* __restrict__ ds, const uint32_t* __restrict__ dcell,
  const typecode* __restrict__ code, uint32_t dcc, uint32_t* __restrict__ cellpart, uint32_t* __restrict__ sortpart)
{
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=ds->nptot)return;
  const uint32_t rcell=dcell[p];
  const uint32_t cx=B2S_DCEL_CELLX(dcc,rcell)-ds->cellmin[0];
  const uint32_t cy=B2S_DCEL_CELLY(dcc,rcell)-ds->cellmin[1];
  const uint32_t cz=B2S_DCEL_CELLZ(dcc,rcell)-ds->cellmin[2];
  const uint32_t cellsort=cx+cy*ds->ncx+cz*ds->nsheet;
  const typecode rcode=code[p];
  const uint32_t codetype=B2S_CODE_TYPE(rcode);
  const uint32_t codeout=B2S_CODE_SPECIAL(rcode);
  uint32_t box;
  if(codetype<B2S_CODE_TYPE_FLOATING){
    box=(codeout<B2S_CODE_OUTIGNORE? ((cx<ds->ncx && cy<ds->ncy && cz<ds->ncz)? cellsort: ds->box_boundignore)
        : (codeout==B2S_CODE_OUTIGNORE? ds->box_boundoutignore: ds->box_boundout));
  }
  else{
    box=(codeout<=B2S_CODE_OUTIGNORE? (codeout<B2S_CODE_OUTIGNORE? ds->box_fluid+cellsort: ds->box_fluidoutignore)
        : (codetype==B2S_CODE_TYPE_FLOATING? ds->box_boundout: ds->box_fluidout));
  }
  cellpart[p]=box;
  sortpart[p]=p;
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort, 8-bit digits, (key=cellpart, value=sortpart).  The permutation of a stable
// sort is unique, so it equals the reference's counting sort (JCellDivCpuSingle.cpp:186-238) and
// thrust::stable_sort_by_key (-stable) bit for bit.
//   pass = k_radix_hist (tile digit histograms) -> k_scan (exclusive, digit-major) -> k_radix_scatter.
// Traffic per pass: 4 B (hist) + 8 B in + 8 B out per particle; 2-3 passes cover log2(2*Nct+5) bits.
// ------------------------------------------------------------------------------------------------
#define RDX_ITEMS 16
#define RDX_TILE (DIV_BS*RDX_ITEMS)

// NBINS = 256 or 512: the passes split ceil(log2(maxkey)) bits as evenly as they can, so the ~25 key bits of a 10M-particle dam
// break (2*Nct+5 boxes) take three passes of 9+8+8 bits instead of four of 8 (radix_plan).
template<int NBINS> __global__ void __launch_bounds__(DIV_BS) k_radix_hist(const uint32_t* __restrict__ keys, const DivState* __restrict__ ds,
  int shift, uint32_t ntiles, uint32_t* __restrict__ counts)
{
  __shared__ uint32_t h[NBINS];
  const uint32_t n=ds->nptot;
  #pragma unroll
  for(int b=threadIdx.x;b<NBINS;b+=DIV_BS)h[b]=0;
  __syncthreads();
  const uint32_t base=blockIdx.x*RDX_TILE;
  #pragma unroll
  for(int r=0;r<RDX_ITEMS;r++){
    const uint32_t i=base+r*DIV_BS+threadIdx.x;
    if(i<n)atomicAdd(&h[(keys[i]>>shift)&uint32_t(NBINS-1)],1u);
  }
  __syncthreads();
  #pragma unroll
  for(int b=threadIdx.x;b<NBINS;b+=DIV_BS)counts[size_t(b)*ntiles+blockIdx.x]=h[b];
}

// Exclusive scan of n uints (n = 256*ntiles, ~6e5 at 10M particles).  Blocks of 4096 elements:
//   k_scan_partial (block totals) -> k_scan_excl on the totals (single block, tiny) -> k_scan_final (local scan + block offset).
// Small inputs (n <= 4096*4) use the single-block kernel alone.
#define SCAN_BLK 4096
__device__ __forceinline__ uint32_t block_excl_scan4(uint32_t (&v)[4], uint32_t* wsum, uint32_t& total){
  // each of the 1024 threads owns 4 consecutive elements; returns the exclusive prefix of v[0] inside the block
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  const uint32_t tsum=v[0]+v[1]+v[2]+v[3];
  uint32_t x=tsum;
  #pragma unroll
  for(int o=1;o<32;o<<=1){ const uint32_t y=__shfl_up_sync(0xffffffffu,x,o); if(lane>=o)x+=y; }
  if(lane==31)wsum[w]=x;
  __syncthreads();
  if(w==0){
    uint32_t s=wsum[lane];
    #pragma unroll
    for(int o=1;o<32;o<<=1){ const uint32_t y=__shfl_up_sync(0xffffffffu,s,o); if(lane>=o)s+=y; }
    wsum[lane]=s;
  }
  __syncthreads();
  total=wsum[31];
  return (w? wsum[w-1]: 0u)+(x-tsum);
}
__global__ void __launch_bounds__(1024) k_scan_excl(uint32_t* data, uint32_t n){
  __shared__ uint32_t wsum[32];
  uint32_t carry=0;
  for(uint32_t base=0; base<n; base+=SCAN_BLK){
    const uint32_t i0=base+threadIdx.x*4;
    uint32_t v[4];
    #pragma unroll
    for(int k=0;k<4;k++)v[k]=(i0+k<n? data[i0+k]: 0u);
    uint32_t total;
    uint32_t excl=carry+block_excl_scan4(v,wsum,total);
    #pragma unroll
    for(int k=0;k<4;k++){ if(i0+k<n)data[i0+k]=excl; excl+=v[k]; }
    carry+=total;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(1024) k_scan_partial(const uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ bsum){
  __shared__ uint32_t wsum[32];
  const uint32_t i0=blockIdx.x*SCAN_BLK+threadIdx.x*4;
  uint32_t t=0;
  if(i0+3<n){ const uint4 q=*reinterpret_cast<const uint4*>(data+i0); t=q.x+q.y+q.z+q.w; }
  else for(int k=0;k<4;k++)if(i0+k<n)t+=data[i0+k];
  t=__reduce_add_sync(0xffffffffu,t);
  if((threadIdx.x&31)==0)wsum[threadIdx.x>>5]=t;
  __syncthreads();
  if(threadIdx.x<32){ t=__reduce_add_sync(0xffffffffu,wsum[threadIdx.x]); if(threadIdx.x==0)bsum[blockIdx.x]=t; }
}
__global__ void __launch_bounds__(1024) k_scan_final(uint32_t* data, uint32_t n, const uint32_t* __restrict__ bsum_excl){
  __shared__ uint32_t wsum[32];
  const uint32_t i0=blockIdx.x*SCAN_BLK+threadIdx.x*4;
  uint32_t v[4];
  #pragma unroll
  for(int k=0;k<4;k++)v[k]=(i0+k<n? data[i0+k]: 0u);
  uint32_t total;
  uint32_t excl=bsum_excl[blockIdx.x]+block_excl_scan4(v,wsum,total);
  #pragma unroll
  for(int k=0;k<4;k++){ if(i0+k<n)data[i0+k]=excl; excl+=v[k]; }
}
// scratch: at least n/SCAN_BLK+1 uints (may be NULL for n <= 4*SCAN_BLK).  Returns the number of launches.
int b2s_scan_excl(cudaStream_t s, uint32_t* data, uint32_t n, uint32_t* scratch){
  if(n<=4u*SCAN_BLK || !scratch){ k_scan_excl<<<1,1024,0,s>>>(data,n); return 1; }
  const uint32_t nb=(n+SCAN_BLK-1)/SCAN_BLK;
  k_scan_partial<<<nb,1024,0,s>>>(data,n,scratch);
  k_scan_excl<<<1,1024,0,s>>>(scratch,nb);
  k_scan_final<<<nb,1024,0,s>>>(data,n,scratch);
  return 3;
}

template<int NBINS,int PRE> __global__ void __launch_bounds__(DIV_BS) k_radix_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ offsets,
  const DivState* __restrict__ ds, int shift, uint32_t ntiles)
{
  __shared__ uint32_t s_base[NBINS];
  __shared__ uint32_t s_wc[DIV_BS/32][NBINS];
  const uint32_t n=ds->nptot;
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  #pragma unroll
  for(int b=threadIdx.x;b<NBINS;b+=DIV_BS)s_base[b]=offsets[size_t(b)*ntiles+blockIdx.x];
  const uint32_t base=blockIdx.x*RDX_TILE;
  // PRE keys and values at a time into registers first: 2*PRE independent loads in flight per thread instead of two between barriers
  uint32_t rk[PRE], rv[PRE];
  #pragma unroll
  for(int r=0;r<RDX_ITEMS;r++){
    if(r%PRE==0){
      #pragma unroll
      for(int k=0;k<PRE;k++){ const uint32_t i2=base+(r+k)*DIV_BS+threadIdx.x; rk[k]=(i2<n? keys_in[i2]: 0u); rv[k]=(i2<n? vals_in[i2]: 0u); }
    }
    const uint32_t i=base+r*DIV_BS+threadIdx.x;
    const bool valid=(i<n);
    const uint32_t key=rk[r%PRE];
    const uint32_t val=rv[r%PRE];
    const uint32_t digit=(key>>shift)&uint32_t(NBINS-1);
    #pragma unroll
    for(int b=threadIdx.x;b<NBINS;b+=DIV_BS){
      #pragma unroll
      for(int k=0;k<DIV_BS/32;k++)s_wc[k][b]=0;
    }
    __syncthreads();
    const uint32_t peers=__match_any_sync(0xffffffffu,(valid? digit: (uint32_t(NBINS)+lane)));
    const uint32_t lower=peers&((1u<<lane)-1u);
    if(valid && lower==0)s_wc[w][digit]=__popc(peers);
    __syncthreads();
    if(valid){
      uint32_t pre=0;
      for(int k=0;k<w;k++)pre+=s_wc[k][digit];
      const uint32_t pos=s_base[digit]+pre+__popc(lower);
      keys_out[pos]=key; vals_out[pos]=val;
    }
    __syncthreads();
    #pragma unroll
    for(int b=threadIdx.x;b<NBINS;b+=DIV_BS){
      uint32_t tot=0;
      #pragma unroll
      for(int k=0;k<DIV_BS/32;k++)tot+=s_wc[k][b];
      s_base[b]+=tot;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Begin table (CalcBeginEndCell, JCellDivGpu_ker.cu:533-570).  Stored in the CPU form
// begincell[box] = first particle of box, begincell[nctt-1] = nptot (JCellDivCpu.h BeginCell), which
// makes a row of x-contiguous cells one [begin[c0], begin[c1+1]) range (JCellSearch_iker.h:155-167).
// ------------------------------------------------------------------------------------------------
// One thread per table entry: begin[c] = number of keys below c = lower bound of c in the sorted keys (24 probes at 10M keys,
// all inside L2).  The former one-thread-per-particle form filled the runs of EMPTY boxes serially from the particle in front
// of them, and a dam-break domain is mostly empty boxes (190 us at 10M particles for a 40 MB read).
__global__ void __launch_bounds__(DIV_BS) k_beginend(const DivState* __restrict__ ds, const uint32_t* __restrict__ keys, uint32_t* __restrict__ beg){
  // grid-stride: the table size is a device-side value (the cell domain of THIS divide); the host only knows the map-wide bound,
  // which is many times larger on a slab rank
  const uint32_t nctt=ds->nctt, n=ds->nptot;
  for(uint32_t c=blockIdx.x*blockDim.x+threadIdx.x; c<nctt; c+=gridDim.x*blockDim.x){
    uint32_t lo=0, hi=n;
    while(lo<hi){ const uint32_t mid=lo+((hi-lo)>>1); if(keys[mid]<c)lo=mid+1; else hi=mid; }
    beg[c]=lo;
  }
}

// Cells with a boundary particle within +-d cells (the neighbour window of the interaction): every occupied boundary cell marks
// its window.  Lets the fluid interaction skip its whole boundary pass for the warps far from any wall (forces_pair2.cu).
__global__ void __launch_bounds__(DIV_BS) k_nearbound_clear(const DivState* __restrict__ ds, uint32_t* __restrict__ nearb4){
  const uint32_t n4=(ds->nct+3u)>>2;
  for(uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; i<n4; i+=gridDim.x*blockDim.x)nearb4[i]=0u;
}
__global__ void __launch_bounds__(DIV_BS) k_nearbound_mark(const DivState* __restrict__ ds, const uint32_t* __restrict__ beg, uint8_t* __restrict__ nearb, int d){
  const int ncx=int(ds->ncx), ncy=int(ds->ncy), ncz=int(ds->ncz), nsheet=int(ds->nsheet);
  const uint32_t nct=ds->nct;
  for(uint32_t c=blockIdx.x*blockDim.x+threadIdx.x; c<nct; c+=gridDim.x*blockDim.x){
    if(beg[c+1]==beg[c])continue;
    const int cz=int(c)/nsheet, cy=(int(c)-cz*nsheet)/ncx, cx=int(c)-cz*nsheet-cy*ncx;
    for(int z=max(cz-d,0);z<=min(cz+d,ncz-1);z++)for(int y=max(cy-d,0);y<=min(cy+d,ncy-1);y++)for(int x=max(cx-d,0);x<=min(cx+d,ncx-1);x++)nearb[x+y*ncx+z*nsheet]=1;
  }
}

// Particle counts after the divide (JCellDivGpuSingle.cpp:218-231), one thread.
__global__ void k_divcounts(DivState* ds, const uint32_t* __restrict__ beg){
  if(threadIdx.x||blockIdx.x)return;
  const uint32_t npb_in=ds->npb;
  #define CSIZE(b) (beg[(b)+1]-beg[(b)])
  ds->npb_ignore=CSIZE(ds->box_boundignore);
  ds->npb_out=CSIZE(ds->box_boundout);
  ds->npf_out=CSIZE(ds->box_fluidout);
  ds->npb_outignore=CSIZE(ds->box_boundoutignore);
  ds->npf_outignore=CSIZE(ds->box_fluidoutignore);
  #undef CSIZE
  ds->np=ds->nptot-ds->npb_out-ds->npf_out-ds->npb_outignore-ds->npf_outignore;
  ds->npb=npb_in-ds->npb_outignore;
  ds->npbok=ds->npb-ds->npb_ignore;
  ds->npf_out_total+=ds->npf_out;
  if(ds->npb_out)ds->error=1;   // JSphGpuSingle::AbortBoundOut (JSphGpuSingle.cpp:505)
}

// ------------------------------------------------------------------------------------------------
// Fused gather: every persistent array through sortpart, + cell-relative position and pressure for
// the interaction (UpdatePosCell + the CPU path's Press_c, JSphCpu.cpp:537-547), + |v|^2 max
// (PreInteraction_Forces, JSphGpu.cpp:898-903), + ridp for moving/floating ids (CalcRidp).
// Algorithmic traffic: 4 (sortpart) + 2x(4+2+4+16+8+16)=100 B + 16 B (posp) per particle.
// ------------------------------------------------------------------------------------------------
struct ReorderArgs {
  PartSet src, dst;
  const double2* pre_xy_s; const double* pre_z_s; const float4* pre_v_s;
  double2* pre_xy_d; double* pre_z_d; float4* pre_v_d;
  const float4* m1_s; float4* m1_d;
  const float* tau_s; float* tau_d;
  const float4* nor_s; float4* nor_d;
  const float4* mace_s; float4* mace_d;
  float4* posp; uint4* posh; uint32_t* ridp;
  uint32_t casenfixed, casenbound;
};


// PPT particles per thread (independent gathers in flight), PAIRH: the FP16 copy of a PAIR of particles is one 16-byte store by its
// even lane (partner values through a shuffle) instead of three 2-byte stores per particle into a sector that is never filled.
template<int PPT,bool PAIRH> __global__ void __launch_bounds__(DIV_BS) k_reorder(const DivState* __restrict__ ds, StepState* st, const uint32_t* __restrict__ sortpart,
  ReorderArgs a, const __grid_constant__ DevCte cte)
{
  const uint32_t n=ds->nptot;
  float v2=0.f;
  #pragma unroll
  for(int it=0;it<PPT;it++){
    const uint32_t p=(blockIdx.x*PPT+it)*blockDim.x+threadIdx.x;
    uint32_t hx=0,hy=0,hz=0; bool live=false;
    if(p<n){
      const uint32_t s=sortpart[p];
      const uint32_t id=a.src.idp[s];
      const typecode cod=a.src.code[s];
      const uint32_t dc=a.src.dcell[s];
      const double2 xy=a.src.posxy[s];
      const double z=a.src.posz[s];
      const float4 vr=a.src.velrho[s];
      a.dst.idp[p]=id; a.dst.code[p]=cod; a.dst.dcell[p]=dc; a.dst.posxy[p]=xy; a.dst.posz[p]=z; a.dst.velrho[p]=vr;
      if(a.pre_v_s){ a.pre_xy_d[p]=a.pre_xy_s[s]; a.pre_z_d[p]=a.pre_z_s[s]; a.pre_v_d[p]=a.pre_v_s[s]; }
      if(a.m1_s)a.m1_d[p]=a.m1_s[s];
      if(a.tau_s){
        const float2* ts=(const float2*)(a.tau_s+size_t(s)*6); float2* td=(float2*)(a.tau_d+size_t(p)*6);
        td[0]=ts[0]; td[1]=ts[1]; td[2]=ts[2];
      }
      if(a.nor_s)a.nor_d[p]=a.nor_s[s];
      if(a.mace_s)a.mace_d[p]=a.mace_s[s];
      if(p<ds->np){
        // position relative to the origin of the particle's own cell (x: of its block of cells) (PosCell idea, JSphGpuSimple_ker.cu:46-57)
        const uint32_t cx=B2S_DCEL_CELLX(cte.domcellcode,dc), cy=B2S_DCEL_CELLY(cte.domcellcode,dc), cz=B2S_DCEL_CELLZ(cte.domcellcode,dc);
        float4 pp;
        pp.x=float((xy.x-cte.domposmin[0])-cte.scell_d*double((cx>>B2S_XBLK)<<B2S_XBLK));   // block-relative x (b2s_internal.h)
        pp.y=float((xy.y-cte.domposmin[1])-cte.scell_d*double(cy));
        pp.z=float((z   -cte.domposmin[2])-cte.scell_d*double(cz));
        pp.w=b2s_press(vr.w,cte);
        a.posp[p]=pp;
        { // FP16 pre-filter copy in cell units (pair-interleaved, see b2s_internal.h posh)
          const float ovs=1.0f/cte.scell;
          const __half h0=__float2half_rn(pp.x*ovs), h1=__float2half_rn(pp.y*ovs), h2=__float2half_rn(pp.z*ovs);
          if(PAIRH){ hx=__half_as_ushort(h0); hy=__half_as_ushort(h1); hz=__half_as_ushort(h2); live=true; }
          else{ __half* h=reinterpret_cast<__half*>(a.posh+(p>>1))+(p&1u); h[0]=h0; h[2]=h1; h[4]=h2; }
        }
        if((p>=ds->npb || cte.dt_allparticles) && B2S_CODE_ISNORMAL(cod))v2=fmaxf(v2,vr.x*vr.x+vr.y*vr.y+vr.z*vr.z);   // ghosts (slab halos) excluded
        if(a.ridp && id>=a.casenfixed && id<a.casenbound && B2S_CODE_ISNORMAL(cod))a.ridp[id-a.casenfixed]=p;   // never a duplicate (CalcRidp with PeriActive, JSphCpu.cpp)
      }
    }
    if(PAIRH){
      // lanes (2k, 2k+1) hold particles (p, p+1) of one 16-byte record (p is even for even lanes: blockDim is a multiple of 32)
      const uint32_t ox=__shfl_down_sync(0xffffffffu,hx,1), oy=__shfl_down_sync(0xffffffffu,hy,1), oz=__shfl_down_sync(0xffffffffu,hz,1);
      const bool olive=__shfl_down_sync(0xffffffffu,live? 1: 0,1)!=0;
      if(live && !(threadIdx.x&1))a.posh[p>>1]=make_uint4(hx|((olive? ox: 0u)<<16),hy|((olive? oy: 0u)<<16),hz|((olive? oz: 0u)<<16),0u);
    }
  }
  // block max -> one atomic per warp (values >=0 so uint order == float order)
  #pragma unroll
  for(int o=16;o;o>>=1)v2=fmaxf(v2,__shfl_xor_sync(0xffffffffu,v2,o));
  if((threadIdx.x&31)==0 && v2>0.f)atomicMax(&st->velmax2_bits,__float_as_uint(v2));
}

// Excluded fluid particles of this divide: after the sort they sit right behind the valid ones, [np+npb_out, +npf_out).
// Record: u32 idp | u32 motive (1 position, 2 density, 3 displacement: JPartOutBi4Save) | f64 pos[3] | f32 vel[3] | f32 rho.
struct OutRec { uint32_t idp, motive; double pos[3]; float vel[3]; float rho; };
__global__ void __launch_bounds__(DIV_BS) k_collect_out(const DivState* __restrict__ ds, PartSet a, OutRec* out, uint32_t* count, uint32_t cap){
  const uint32_t n=ds->npf_out, p0=ds->np+ds->npb_out;
  for(uint32_t i=blockIdx.x*blockDim.x+threadIdx.x; i<n; i+=gridDim.x*blockDim.x){
    const uint32_t p=p0+i;
    const uint32_t k=atomicAdd(count,1u);
    if(k>=cap)continue;   // overflow: the count keeps growing and b2s_out_fetch reports it
    const typecode c=a.code[p];
    const uint32_t sp=B2S_CODE_SPECIAL(c);
    OutRec r; r.idp=a.idp[p]; r.motive=(sp==B2S_CODE_OUTRHO? 2u: (sp==B2S_CODE_OUTMOV? 3u: 1u));
    const double2 xy=a.posxy[p]; r.pos[0]=xy.x; r.pos[1]=xy.y; r.pos[2]=a.posz[p];
    const float4 v=a.velrho[p]; r.vel[0]=v.x; r.vel[1]=v.y; r.vel[2]=v.z; r.rho=v.w;
    out[k]=r;
  }
}

__global__ void k_reset_step_maxima(StepState* st){
  st->velmax2_bits=0; st->acemax2_bits=0; st->viscdtmax_bits=0;
}

// ------------------------------------------------------------------------------------------------
// Passes of the LSD sort over the bits of the largest key: as few as 9-bit digits allow, the bits spread evenly (8 or 9 per pass).
static inline int radix_plan(uint32_t maxkey,int* digitbits){
  int bits=1; while(bits<32 && (maxkey>>bits)!=0)bits++;
  const int passes=(bits+8)/9;
  for(int k=0;k<passes;k++)digitbits[k]=std::max(8,(bits+passes-1-k)/passes);   // e.g. 25 bits -> 9,8,8
  return passes;
}

int b2s_celldiv_impl(b2s_ctx* ctx){
  cudaStream_t s=ctx->stream;
  const uint32_t nptot=ctx->np_host;      // upper bound known to the host; == device np (outs already dropped)
  const uint32_t npb=ctx->npb_host;
  const DevCte& c=ctx->cte;
  if(nptot==0)return B2S_OK;
  // 1. limits
  const bool mixed=(ctx->slab_on || c.periactive);   // ghosts / periodic duplicates are appended behind the owned particles: limits by type
  k_limits_reset<<<1,32,0,s>>>(ctx->div,(ctx->bound_changed || mixed)? 1: 0); B2S_LAUNCHED(ctx);
  const unsigned gl=min(b2s_blocks(nptot,DIV_BS),148u*8u);
  if(mixed){ k_limits_bytype<<<gl,DIV_BS,0,s>>>(ctx->div,ctx->cur.dcell,ctx->cur.code,c.domcellcode); B2S_LAUNCHED(ctx); }
  else if(ctx->bound_changed && npb){ k_limits<<<gl,DIV_BS,0,s>>>(npb,0,ctx->cur.dcell,ctx->cur.code,c.domcellcode,ctx->div->blim); B2S_LAUNCHED(ctx); }
  if(!mixed && nptot>npb){ k_limits<<<gl,DIV_BS,0,s>>>(nptot-npb,npb,ctx->cur.dcell,ctx->cur.code,c.domcellcode,ctx->div->flim); B2S_LAUNCHED(ctx); }
  k_divstate<<<1,32,0,s>>>(ctx->div,c.domcells[0],c.domcells[1],c.domcells[2],(uint32_t)c.scelldiv); B2S_LAUNCHED(ctx);
  // 2. keys
  k_presort<<<b2s_blocks(nptot,DIV_BS),DIV_BS,0,s>>>(ctx->div,ctx->cur.dcell,ctx->cur.code,c.domcellcode,ctx->cellpart,ctx->sortpart); B2S_LAUNCHED(ctx);
  // 3. stable radix sort
  const uint32_t ntiles=b2s_blocks(nptot,RDX_TILE);
  static int rpre=-1;     // B2S_RADIX_PRE=1|4|8|16: keys preloaded per batch in the scatter (experiment switch, tools/gpu_job.sh radix_ab)
  if(rpre<0){ const char* e=getenv("B2S_RADIX_PRE"); rpre=(e? atoi(e): 8); }   // profiles/r02y_radix_ab.txt: 1: 20.62, 4: 20.76, 8: 20.58, 16: 20.56 ms per step
  int dbits[4];
  const int passes=radix_plan(2u*ctx->nct_max+5u,dbits);
  uint32_t *kin=ctx->cellpart,*vin=ctx->sortpart,*kout=ctx->cellpart_alt,*vout=ctx->sortpart_alt;
  for(int ps=0,shift=0;ps<passes;shift+=dbits[ps],ps++){
    if(dbits[ps]>8){
      k_radix_hist<512><<<ntiles,DIV_BS,0,s>>>(kin,ctx->div,shift,ntiles,ctx->radix_counts); B2S_LAUNCHED(ctx);
      ctx->launches+=b2s_scan_excl(s,ctx->radix_counts,512u*ntiles,ctx->radix_counts+ctx->radix_counts_len);
      if(rpre==16)k_radix_scatter<512,16><<<ntiles,DIV_BS,0,s>>>(kin,vin,kout,vout,ctx->radix_counts,ctx->div,shift,ntiles);
      else if(rpre==8)k_radix_scatter<512,8><<<ntiles,DIV_BS,0,s>>>(kin,vin,kout,vout,ctx->radix_counts,ctx->div,shift,ntiles);
      else if(rpre==4)k_radix_scatter<512,4><<<ntiles,DIV_BS,0,s>>>(kin,vin,kout,vout,ctx->radix_counts,ctx->div,shift,ntiles);
      else k_radix_scatter<512,1><<<ntiles,DIV_BS,0,s>>>(kin,vin,kout,vout,ctx->radix_counts,ctx->div,shift,ntiles);
      B2S_LAUNCHED(ctx);
    }
    else{
      k_radix_hist<256><<<ntiles,DIV_BS,0,s>>>(kin,ctx->div,shift,ntiles,ctx->radix_counts); B2S_LAUNCHED(ctx);
      ctx->launches+=b2s_scan_excl(s,ctx->radix_counts,256u*ntiles,ctx->radix_counts+ctx->radix_counts_len);
      if(rpre==16)k_radix_scatter<256,16><<<ntiles,DIV_BS,0,s>>>(kin,vin,kout,vout,ctx->radix_counts,ctx->div,shift,ntiles);
      else if(rpre==8)k_radix_scatter<256,8><<<ntiles,DIV_BS,0,s>>>(kin,vin,kout,vout,ctx->radix_counts,ctx->div,shift,ntiles);
      else if(rpre==4)k_radix_scatter<256,4><<<ntiles,DIV_BS,0,s>>>(kin,vin,kout,vout,ctx->radix_counts,ctx->div,shift,ntiles);
      else k_radix_scatter<256,1><<<ntiles,DIV_BS,0,s>>>(kin,vin,kout,vout,ctx->radix_counts,ctx->div,shift,ntiles);
      B2S_LAUNCHED(ctx);
    }
    uint32_t* t=kin; kin=kout; kout=t; t=vin; vin=vout; vout=t;
  }
  if(kin!=ctx->cellpart){ // odd number of passes: sorted data sits in the alt buffers -> swap names
    uint32_t* t=ctx->cellpart; ctx->cellpart=ctx->cellpart_alt; ctx->cellpart_alt=t;
    t=ctx->sortpart; ctx->sortpart=ctx->sortpart_alt; ctx->sortpart_alt=t;
  }
  // 4. begin table + counts
  k_beginend<<<min(b2s_blocks(ctx->nctt_max,DIV_BS),148u*16u),DIV_BS,0,s>>>(ctx->div,ctx->cellpart,(uint32_t*)ctx->begincell); B2S_LAUNCHED(ctx);
  k_divcounts<<<1,32,0,s>>>(ctx->div,(const uint32_t*)ctx->begincell); B2S_LAUNCHED(ctx);
  ctx->nearbound_valid=false;
  if(ctx->npb_host){
    ctx->nearbound_valid=true;
    const unsigned gn=min(b2s_blocks(ctx->nct_max,DIV_BS),148u*16u);
    k_nearbound_clear<<<gn,DIV_BS,0,s>>>(ctx->div,(uint32_t*)ctx->nearbound); B2S_LAUNCHED(ctx);
    k_nearbound_mark<<<gn,DIV_BS,0,s>>>(ctx->div,(const uint32_t*)ctx->begincell,ctx->nearbound,c.scelldiv); B2S_LAUNCHED(ctx);
  }
  // 5. fused gather
  k_reset_step_maxima<<<1,1,0,s>>>(ctx->step); B2S_LAUNCHED(ctx);
  ReorderArgs a; memset(&a,0,sizeof(a));
  a.src=ctx->cur; a.dst=ctx->alt;
  if(ctx->pre_active){
    a.pre_xy_s=ctx->pospre_xy; a.pre_z_s=ctx->pospre_z; a.pre_v_s=ctx->velrhopre;
    a.pre_xy_d=ctx->pospre_xy_alt; a.pre_z_d=ctx->pospre_z_alt; a.pre_v_d=ctx->velrhopre_alt;
  }
  if(ctx->velrhom1){ a.m1_s=ctx->velrhom1; a.m1_d=ctx->velrhom1_alt; }
  if(ctx->spstau){ a.tau_s=ctx->spstau; a.tau_d=ctx->spstau_alt; }
  if(ctx->boundnor){ a.nor_s=ctx->boundnor; a.nor_d=ctx->boundnor_alt; }
  if(ctx->motionace){ a.mace_s=ctx->motionace; a.mace_d=ctx->motionace_alt; }
  a.posp=ctx->posp; a.posh=ctx->posh;
  a.ridp=(ctx->prm.casenmoving+ctx->prm.casenfloat? ctx->ridp: NULL);
  a.casenfixed=ctx->prm.casenfixed; a.casenbound=ctx->prm.casenbound;
  if(a.ridp)B2S_CUDA(ctx,cudaMemsetAsync(ctx->ridp,0xff,size_t(ctx->prm.casenmoving+ctx->prm.casenfloat)*4,s));
  { // B2S_REORDER=<particles per thread 1|2><pair store 0|1>[,threads]: experiment switch of tools/gpu_job.sh reorder_ab
    static int mode=-1, bs=DIV_BS;
    // measured on the B200 at 10M particles (profiles/r02w_reorder_ab.txt, ms per step): 10: 21.18, 11: 21.22, 20: 20.71, 21: 20.84,
    // 128-thread blocks: worse; profiles/r02x_reorder_ab.txt: 40: 20.64).  Four particles per thread (more independent gathers in
    // flight) is the default.
    if(mode<0){ const char* e=getenv("B2S_REORDER"); mode=(e? atoi(e): 40); if(e && strchr(e,','))bs=atoi(strchr(e,',')+1); if(bs<32||bs>DIV_BS||(bs&31))bs=DIV_BS; }
    const int ppt=(mode/10==4? 4: (mode/10==2? 2: 1)); const bool pairh=(mode%10)!=0;
    const unsigned gb=b2s_blocks(nptot,unsigned(bs*ppt));
    if(ppt==4)k_reorder<4,false><<<gb,bs,0,s>>>(ctx->div,ctx->step,ctx->sortpart,a,ctx->cte);
    else if(ppt==2){ if(pairh)k_reorder<2,true><<<gb,bs,0,s>>>(ctx->div,ctx->step,ctx->sortpart,a,ctx->cte); else k_reorder<2,false><<<gb,bs,0,s>>>(ctx->div,ctx->step,ctx->sortpart,a,ctx->cte); }
    else{ if(pairh)k_reorder<1,true><<<gb,bs,0,s>>>(ctx->div,ctx->step,ctx->sortpart,a,ctx->cte); else k_reorder<1,false><<<gb,bs,0,s>>>(ctx->div,ctx->step,ctx->sortpart,a,ctx->cte); }
    B2S_LAUNCHED(ctx);
  }
  { PartSet t=ctx->cur; ctx->cur=ctx->alt; ctx->alt=t; }
  if(ctx->out_dev){ k_collect_out<<<32,DIV_BS,0,s>>>(ctx->div,ctx->cur,(OutRec*)ctx->out_dev,ctx->out_count_dev,ctx->out_cap); B2S_LAUNCHED(ctx); }
  if(ctx->pre_active){
    std::swap(ctx->pospre_xy,ctx->pospre_xy_alt); std::swap(ctx->pospre_z,ctx->pospre_z_alt); std::swap(ctx->velrhopre,ctx->velrhopre_alt);
  }
  if(ctx->velrhom1)std::swap(ctx->velrhom1,ctx->velrhom1_alt);
  if(ctx->spstau)std::swap(ctx->spstau,ctx->spstau_alt);
  if(ctx->boundnor)std::swap(ctx->boundnor,ctx->boundnor_alt);
  if(ctx->motionace)std::swap(ctx->motionace,ctx->motionace_alt);
  ctx->bound_changed=false;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}

// ---- excluded particles: enable the collection / drain it ----------------------------------------------------
extern "C" int b2s_out_enable(b2s_ctx* ctx, uint32_t capacity){
  if(!ctx || !capacity)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  if(ctx->out_dev){ cudaFree(ctx->out_dev); cudaFree(ctx->out_count_dev); ctx->out_dev=NULL; ctx->out_count_dev=NULL; }
  B2S_CUDA(ctx,cudaMalloc(&ctx->out_dev,size_t(capacity)*sizeof(OutRec)));
  B2S_CUDA(ctx,cudaMalloc((void**)&ctx->out_count_dev,4));
  B2S_CUDA(ctx,cudaMemsetAsync(ctx->out_count_dev,0,4,ctx->stream));
  ctx->out_cap=capacity;
  return B2S_OK;
}
extern "C" int b2s_out_fetch(b2s_ctx* ctx, uint32_t maxcount, uint32_t* count_out, uint32_t* idp, double* pos, float* vel, float* rhop, uint8_t* motive){
  if(!ctx || !count_out)return B2S_ERR_ARG;
  *count_out=0;
  if(!ctx->out_dev)return b2s_fail(ctx,B2S_ERR_STATE,"excluded particles are not collected (b2s_out_enable)");
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  uint32_t n=0;
  B2S_CUDA(ctx,cudaMemcpyAsync(&n,ctx->out_count_dev,4,cudaMemcpyDeviceToHost,ctx->stream));
  B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
  if(n>ctx->out_cap)return b2s_fail(ctx,B2S_ERR_NOMEM,"%u particles were excluded since the last fetch but the buffer holds %u",n,ctx->out_cap);
  if(n>maxcount)return b2s_fail(ctx,B2S_ERR_ARG,"b2s_out_fetch: %u records pending, room for %u",n,maxcount);
  if(n){
    std::vector<OutRec> h(n);
    B2S_CUDA(ctx,cudaMemcpyAsync(h.data(),ctx->out_dev,size_t(n)*sizeof(OutRec),cudaMemcpyDeviceToHost,ctx->stream));
    B2S_CUDA(ctx,cudaMemsetAsync(ctx->out_count_dev,0,4,ctx->stream));
    B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream));
    for(uint32_t i=0;i<n;i++){
      if(idp)idp[i]=h[i].idp;
      if(pos){ pos[size_t(i)*3]=h[i].pos[0]; pos[size_t(i)*3+1]=h[i].pos[1]; pos[size_t(i)*3+2]=h[i].pos[2]; }
      if(vel){ vel[size_t(i)*3]=h[i].vel[0]; vel[size_t(i)*3+1]=h[i].vel[1]; vel[size_t(i)*3+2]=h[i].vel[2]; }
      if(rhop)rhop[i]=h[i].rho;
      if(motive)motive[i]=uint8_t(h[i].motive);
    }
  }
  *count_out=n;
  return B2S_OK;
}

void b2s_launch_scan_excl(cudaStream_t s, uint32_t* data, uint32_t n){ k_scan_excl<<<1,1024,0,s>>>(data,n); }   // small inputs (slab.cu)
// nseg independent exclusive scans of n elements each, segment k at data+k*n: one block per segment, one launch (slab.cu: the four
// count arrays of an exchange)
__global__ void __launch_bounds__(1024) k_scan_excl_seg(uint32_t* data, uint32_t n){
  __shared__ uint32_t wsum[32];
  uint32_t* d=data+size_t(blockIdx.x)*n;
  uint32_t carry=0;
  for(uint32_t base=0; base<n; base+=SCAN_BLK){
    const uint32_t i0=base+threadIdx.x*4;
    uint32_t v[4];
    #pragma unroll
    for(int k=0;k<4;k++)v[k]=(i0+k<n? d[i0+k]: 0u);
    uint32_t total;
    uint32_t excl=carry+block_excl_scan4(v,wsum,total);
    #pragma unroll
    for(int k=0;k<4;k++){ if(i0+k<n)d[i0+k]=excl; excl+=v[k]; }
    carry+=total;
    __syncthreads();
  }
}
void b2s_launch_scan_excl_seg(cudaStream_t s, uint32_t* data, uint32_t n, uint32_t nseg){ k_scan_excl_seg<<<nseg,1024,0,s>>>(data,n); }
