// slab.cu -- 1-D slab decomposition along x for several GPUs (one process / one ctx per GPU).
//
// No reference counterpart: DualSPHysics v5.4 is single-GPU (AVAILABLE_MGPU is hard-false, src/source/DualSphDef.h:96-101).
// The mechanism mirrors the reference's periodic-boundary ghosts (JSphGpuSingle::RunPeriodic, src/source/JSphGpuSingle.cpp:276-366;
// KerPeriodicDuplicate*, src/source/JSphGpu_ker.cu:2267-2399): neighbours' particles within one cell layer of a cut are
// appended as CODE_PERIODIC duplicates before the cell divide, take part in the interaction as neighbours, are frozen by
// the integrators and dropped (CODE_OUTIGNORE) by the next divide.  Particles that cross a cut migrate with their full
// state.  All ranks use the SAME global cell grid (domposmin, dcell code), so cell coordinates are bit-identical to the
// single-GPU run.  The transport (NCCL send/recv over NVLink, or device copies in the single-process test) is the
// caller's: this file only packs/unpacks contiguous device buffers -- deterministic order (scan-based compaction).
//
// Record layout (little endian, 8-byte aligned): u32 idp | u16 code | u16 0 | f64 pos[3] | f32 velrho[4]   (48 B)
//   migrants append: [f32 velrhopre[4] f64 pospre[3] pad8] (Symplectic, predictor active) | [f32 velrhom1[4]] (Verlet)
//   both append:     [f32 boundnor[4]] (mDBC) | [f32 tau[6] pad8] (Laminar+SPS); records are multiples of 16 B
#include "b2s_internal.h"
#include "slab_cuts.h"

#define S_BS 256
#define B2S_SLAB_HDR 16        // bytes in front of the records of a device-counted message (k_slab_header)

// Halo width in cell planes: one kernel radius (scelldiv planes).  With mDBC the boundary ghosts must carry the density the
// owner extrapolates at their ghost node; instead of a second exchange after the mDBC pass every rank recomputes it locally,
// which needs the fluid within one more kernel radius of those boundary ghosts: the halo is two radii wide.
static inline uint32_t slab_halo_planes(const b2s_ctx* ctx){ return uint32_t(ctx->prm.scelldiv)*(ctx->prm.tboundary==B2S_BC_MDBC? 2u: 1u); }

void b2s_launch_scan_excl(cudaStream_t s, uint32_t* data, uint32_t n);   // celldiv.cu
void b2s_launch_scan_excl_seg(cudaStream_t s, uint32_t* data, uint32_t n, uint32_t nseg);

struct SlabLayout { uint32_t bytes, off_pre, off_m1, off_tau, off_nor; };

static SlabLayout slab_layout(const b2s_ctx* ctx, int kind){
  SlabLayout l; l.bytes=48; l.off_pre=l.off_m1=l.off_tau=l.off_nor=0;
  // every float4 field sits on a 16-byte boundary and the record size is a multiple of 16 (vector loads/stores)
  if(kind==1 && ctx->pre_active){ l.off_pre=l.bytes; l.bytes+=48; }   // f32 velrhopre[4] | f64 pospre[3] | 8 B pad
  if(kind==1 && ctx->velrhom1){ l.off_m1=l.bytes; l.bytes+=16; }
  if(ctx->boundnor){ l.off_nor=l.bytes; l.bytes+=16; }
  if(ctx->spstau){ l.off_tau=l.bytes; l.bytes+=32; }                  // f32 tau[6] | 8 B pad
  return l;
}

struct SlabArrays {
  uint32_t* idp; typecode* code; uint32_t* dcell; double2* posxy; double* posz; float4* velrho;
  double2* pre_xy; double* pre_z; float4* pre_v; float4* m1; float* tau; float4* nor;
};

// flags: bit0 -> goes to the left neighbour, bit1 -> right neighbour
__global__ void __launch_bounds__(S_BS) k_slab_flag(const DivState* __restrict__ ds, const typecode* __restrict__ code, const uint32_t* __restrict__ dcell,
  uint32_t dcc, int kind, int has_left, int has_right, uint32_t cxa, uint32_t cxb, uint32_t d, uint8_t* flags, uint32_t* counts, uint32_t nblocks)
{
  __shared__ uint32_t cl, cr;
  if(threadIdx.x==0){ cl=0; cr=0; }
  __syncthreads();
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  uint8_t f=0;
  if(p<ds->np){
    const typecode c=code[p];
    if(B2S_CODE_SPECIAL(c)==B2S_CODE_NORMAL){
      const uint32_t cx=B2S_DCEL_CELLX(dcc,dcell[p]);
      if(kind==0){ // halo: owned particles in the cell layers next to a cut
        if(has_left && cx>=cxa && cx<cxa+d)f|=1;
        if(has_right && cx<cxb && cx+d>=cxb)f|=2;
      }
      else{        // migrants: owned particles whose cell left the slab
        if(has_left && cx<cxa)f|=1;
        if(has_right && cx>=cxb)f|=2;
      }
    }
    flags[p]=f;
  }
  const uint32_t bl=__ballot_sync(0xffffffffu,f&1), br=__ballot_sync(0xffffffffu,f&2);
  if((threadIdx.x&31)==0){ if(bl)atomicAdd(&cl,__popc(bl)); if(br)atomicAdd(&cr,__popc(br)); }
  __syncthreads();
  if(threadIdx.x==0){ counts[blockIdx.x]=cl; counts[nblocks+1+blockIdx.x]=cr; }
}

__device__ __forceinline__ void slab_write(char* rec, const SlabLayout& l, const SlabArrays& a, uint32_t p){
  *reinterpret_cast<uint32_t*>(rec)=a.idp[p];
  *reinterpret_cast<uint32_t*>(rec+4)=uint32_t(a.code[p]);
  double* d=reinterpret_cast<double*>(rec+8);
  const double2 xy=a.posxy[p]; d[0]=xy.x; d[1]=xy.y; d[2]=a.posz[p];
  *reinterpret_cast<float4*>(rec+32)=a.velrho[p];
  if(l.off_pre){ double* q=reinterpret_cast<double*>(rec+l.off_pre+16); const double2 pxy=a.pre_xy[p]; q[0]=pxy.x; q[1]=pxy.y; q[2]=a.pre_z[p]; *reinterpret_cast<float4*>(rec+l.off_pre)=a.pre_v[p]; }
  if(l.off_m1)*reinterpret_cast<float4*>(rec+l.off_m1)=a.m1[p];
  if(l.off_tau){ float2* t=reinterpret_cast<float2*>(rec+l.off_tau); const float2* s=reinterpret_cast<const float2*>(a.tau+size_t(p)*6); t[0]=s[0]; t[1]=s[1]; t[2]=s[2]; }
  if(l.off_nor)*reinterpret_cast<float4*>(rec+l.off_nor)=a.nor[p];
}

__global__ void __launch_bounds__(S_BS) k_slab_scatter(const DivState* __restrict__ ds, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ offs,
  uint32_t nblocks, SlabLayout l, SlabArrays a, int kind, char* left, char* right, uint32_t cap, typecode migcode)
{
  __shared__ uint32_t wl[S_BS/32], wr[S_BS/32];
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  const uint8_t f=(p<ds->np? flags[p]: uint8_t(0));
  const uint32_t bl=__ballot_sync(0xffffffffu,f&1), br=__ballot_sync(0xffffffffu,f&2);
  if(lane==0){ wl[w]=__popc(bl); wr[w]=__popc(br); }
  __syncthreads();
  uint32_t pl=0, pr=0;
  for(int k=0;k<w;k++){ pl+=wl[k]; pr+=wr[k]; }
  const uint32_t lm=(1u<<lane)-1u;
  if(f&1){ const uint32_t i=offs[blockIdx.x]+pl+__popc(bl&lm); if(i<cap)slab_write(left+size_t(i)*l.bytes,l,a,p); }
  if(f&2){ const uint32_t i=offs[nblocks+1+blockIdx.x]+pr+__popc(br&lm); if(i<cap)slab_write(right+size_t(i)*l.bytes,l,a,p); }
  // migrated away: dropped by the next divide (OUTIGNORE), or kept as a halo ghost of its new owner (PERIODIC; merged exchange)
  if(kind==1 && f)a.code[p]=typecode(B2S_CODE_SETNORMAL(a.code[p])|migcode);
}

// ---- the same exchange without the host: counts stay on the device (slab_comm.cu) -------------------------------------------
// Message of one side: [u32 nmigrants | u32 nhalo | 8 B pad][migrant records][halo records], sent with a fixed size (the
// capacities), so neither side needs the other's counts on the host.  tot_l / tot_r: device totals of the flag scan.
__global__ void __launch_bounds__(S_BS) k_slab_scatter_dev(const DivState* __restrict__ ds, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ offs,
  uint32_t nblocks, SlabLayout l, SlabArrays a, int kind, char* left, char* right, uint32_t cap,
  const uint32_t* __restrict__ migtot_l, const uint32_t* __restrict__ migtot_r, uint32_t migcap, uint32_t migbytes, typecode migcode)
{
  __shared__ uint32_t wl[S_BS/32], wr[S_BS/32];
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  const uint8_t f=(p<ds->np? flags[p]: uint8_t(0));
  const uint32_t bl=__ballot_sync(0xffffffffu,f&1), br=__ballot_sync(0xffffffffu,f&2);
  if(lane==0){ wl[w]=__popc(bl); wr[w]=__popc(br); }
  __syncthreads();
  uint32_t pl=0, pr=0;
  for(int k=0;k<w;k++){ pl+=wl[k]; pr+=wr[k]; }
  const uint32_t lm=(1u<<lane)-1u;
  // halo records start behind the migrant records of the same message
  const size_t basel=B2S_SLAB_HDR+(migtot_l? size_t(min(*migtot_l,migcap))*migbytes: 0), baser=B2S_SLAB_HDR+(migtot_r? size_t(min(*migtot_r,migcap))*migbytes: 0);
  bool sent=false;
  if(f&1){ const uint32_t i=offs[blockIdx.x]+pl+__popc(bl&lm); if(i<cap){ slab_write(left+basel+size_t(i)*l.bytes,l,a,p); sent=true; } }
  if(f&2){ const uint32_t i=offs[nblocks+1+blockIdx.x]+pr+__popc(br&lm); if(i<cap){ slab_write(right+baser+size_t(i)*l.bytes,l,a,p); sent=true; } }
  if(kind==1 && sent)a.code[p]=typecode(B2S_CODE_SETNORMAL(a.code[p])|migcode);   // a migrant that did not fit stays (the error flag is raised by k_slab_header)
}
// Both kinds in one pass (the per-interaction exchange): flags + block counts of migrants and halo particles ...
__global__ void __launch_bounds__(S_BS) k_slab_flag2(const DivState* __restrict__ ds, const typecode* __restrict__ code, const uint32_t* __restrict__ dcell,
  uint32_t dcc, int has_left, int has_right, uint32_t cxa, uint32_t cxb, uint32_t d, uint8_t* flagsM, uint8_t* flagsH, uint32_t* cM, uint32_t* cH, uint32_t nblocks)
{
  __shared__ uint32_t cnt[4];
  if(threadIdx.x<4)cnt[threadIdx.x]=0;
  __syncthreads();
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  uint8_t fm=0, fh=0;
  if(p<ds->np){
    if(B2S_CODE_SPECIAL(code[p])==B2S_CODE_NORMAL){
      const uint32_t cx=B2S_DCEL_CELLX(dcc,dcell[p]);
      if(has_left){ if(cx<cxa)fm|=1; else if(cx<cxa+d)fh|=1; }
      if(has_right){ if(cx>=cxb)fm|=2; else if(cx+d>=cxb)fh|=2; }
    }
    flagsM[p]=fm; flagsH[p]=fh;
  }
  const uint32_t b0=__ballot_sync(0xffffffffu,fm&1), b1=__ballot_sync(0xffffffffu,fm&2), b2=__ballot_sync(0xffffffffu,fh&1), b3=__ballot_sync(0xffffffffu,fh&2);
  if((threadIdx.x&31)==0){ if(b0)atomicAdd(&cnt[0],__popc(b0)); if(b1)atomicAdd(&cnt[1],__popc(b1)); if(b2)atomicAdd(&cnt[2],__popc(b2)); if(b3)atomicAdd(&cnt[3],__popc(b3)); }
  __syncthreads();
  if(threadIdx.x==0){ cM[blockIdx.x]=cnt[0]; cM[nblocks+1+blockIdx.x]=cnt[1]; cH[blockIdx.x]=cnt[2]; cH[nblocks+1+blockIdx.x]=cnt[3]; }
}
// ... and their records: a migrant is never a halo particle of the same exchange (slabs are wider than the halo), so the two
// writes of a thread never touch the same particle's code.
__global__ void __launch_bounds__(S_BS) k_slab_scatter2(const DivState* __restrict__ ds, const uint8_t* __restrict__ flagsM, const uint8_t* __restrict__ flagsH,
  const uint32_t* __restrict__ cM, const uint32_t* __restrict__ cH, uint32_t nblocks, SlabLayout lM, SlabLayout lH, SlabArrays a,
  char* left, char* right, uint32_t migcap, uint32_t halocap, typecode migcode)
{
  __shared__ uint32_t wc[4][S_BS/32];
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  const bool in=(p<ds->np);
  const uint8_t fm=(in? flagsM[p]: uint8_t(0)), fh=(in? flagsH[p]: uint8_t(0));
  const uint32_t b[4]={__ballot_sync(0xffffffffu,fm&1),__ballot_sync(0xffffffffu,fm&2),__ballot_sync(0xffffffffu,fh&1),__ballot_sync(0xffffffffu,fh&2)};
  if(lane==0){ wc[0][w]=__popc(b[0]); wc[1][w]=__popc(b[1]); wc[2][w]=__popc(b[2]); wc[3][w]=__popc(b[3]); }
  __syncthreads();
  if(!(fm|fh))return;
  const uint32_t lm=(1u<<lane)-1u;
  const uint32_t totl=min(cM[nblocks],migcap), totr=min(cM[2*nblocks+1],migcap);      // migrant records in front of the halo records
  #pragma unroll
  for(int k=0;k<4;k++){
    const bool mig=(k<2), lft=((k&1)==0);
    if(!((mig? fm: fh)&(lft? 1: 2)))continue;
    uint32_t pre=0; for(int q=0;q<w;q++)pre+=wc[k][q];
    const uint32_t* offs=(mig? cM: cH)+(lft? 0u: nblocks+1u);
    const uint32_t i=offs[blockIdx.x]+pre+__popc(b[k]&lm);
    if(i>=(mig? migcap: halocap))continue;                    // does not fit: k_slab_header raises the error flag
    char* base=(lft? left: right)+B2S_SLAB_HDR+(mig? size_t(0): size_t(lft? totl: totr)*lM.bytes);
    slab_write(base+size_t(i)*(mig? lM.bytes: lH.bytes),mig? lM: lH,a,p);
    if(mig)a.code[p]=typecode(B2S_CODE_SETNORMAL(a.code[p])|migcode);
  }
}
__global__ void k_slab_header(StepState* st, const uint32_t* __restrict__ cM, const uint32_t* __restrict__ cH, uint32_t nb, uint32_t migcap, uint32_t halocap,
  char* left, char* right, uint32_t* sent_total, int strict)
{
  if(threadIdx.x||blockIdx.x)return;
  const uint32_t ml=(cM? cM[nb]: 0u), mr=(cM? cM[2*nb+1]: 0u), hl=(cH? cH[nb]: 0u), hr=(cH? cH[2*nb+1]: 0u);
  // strict: what does not fit is an error (a step cannot go on without it); otherwise (re-cut rounds) it travels in the next round
  if(strict && (ml>migcap || mr>migcap || hl>halocap || hr>halocap))st->error=B2S_ERR_NOMEM;
  if(left){ uint32_t* h=reinterpret_cast<uint32_t*>(left); h[0]=min(ml,migcap); h[1]=min(hl,halocap); h[2]=0; h[3]=0; }
  if(right){ uint32_t* h=reinterpret_cast<uint32_t*>(right); h[0]=min(mr,migcap); h[1]=min(hr,halocap); h[2]=0; h[3]=0; }
  if(sent_total){ sent_total[0]=ml+mr; sent_total[1]=hl+hr; sent_total[2]+=ml+mr; }   // [2]: migrants since b2s_slab_comm_init
}

__device__ __forceinline__ void slab_unpack_record(DivState* ds, uint32_t p, const char* rec, const SlabLayout& l, const SlabArrays& a, int kind, const DevCte& c);

__global__ void __launch_bounds__(S_BS) k_slab_unpack_msg(DivState* ds, StepState* st, const char* __restrict__ msg, SlabLayout lM, SlabLayout lH, SlabArrays a,
  uint32_t cap, const __grid_constant__ DevCte c)
{
  const uint32_t* h=reinterpret_cast<const uint32_t*>(msg);
  const uint32_t nm=h[0], nh=h[1];
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=nm+nh)return;
  if(uint64_t(ds->np)+nm+nh>cap){ if(i==0)st->error=B2S_ERR_NOMEM; return; }
  const bool mig=(i<nm);
  const char* rec=msg+B2S_SLAB_HDR+(mig? size_t(i)*lM.bytes: size_t(nm)*lM.bytes+size_t(i-nm)*lH.bytes);
  slab_unpack_record(ds,ds->np+i,rec,mig? lM: lH,a,mig? 1: 0,c);
}
__global__ void k_slab_unpack_msg_commit(DivState* ds, const char* __restrict__ msg, uint32_t cap){
  const uint32_t* h=reinterpret_cast<const uint32_t*>(msg);
  if(uint64_t(ds->np)+h[0]+h[1]<=cap){ ds->np+=h[0]+h[1]; ds->npb+=ds->slab_nbound_tmp; }
  ds->slab_nbound_tmp=0;
}

__global__ void __launch_bounds__(S_BS) k_slab_unpack(DivState* ds, const char* __restrict__ buf, uint32_t n, SlabLayout l, SlabArrays a, int kind,
  const __grid_constant__ DevCte c)
{
  const uint32_t i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=n)return;
  slab_unpack_record(ds,ds->np+i,buf+size_t(i)*l.bytes,l,a,kind,c);
}
__device__ __forceinline__ void slab_unpack_record(DivState* ds, uint32_t p, const char* rec, const SlabLayout& l, const SlabArrays& a, int kind, const DevCte& c)
{
  const uint32_t id=*reinterpret_cast<const uint32_t*>(rec);
  typecode cod=typecode(*reinterpret_cast<const uint32_t*>(rec+4));
  if(kind==0)cod=typecode(B2S_CODE_SETNORMAL(cod)|B2S_CODE_PERIODIC);   // halo ghost
  const double* d=reinterpret_cast<const double*>(rec+8);
  a.idp[p]=id; a.code[p]=cod; a.posxy[p]=make_double2(d[0],d[1]); a.posz[p]=d[2];
  a.velrho[p]=*reinterpret_cast<const float4*>(rec+32);
  a.dcell[p]=B2S_DCEL_CELL(c.domcellcode,uint32_t((d[0]-c.domposmin[0])/c.scell_d),uint32_t((d[1]-c.domposmin[1])/c.scell_d),uint32_t((d[2]-c.domposmin[2])/c.scell_d));
  if(l.off_pre){ const double* q=reinterpret_cast<const double*>(rec+l.off_pre+16); a.pre_xy[p]=make_double2(q[0],q[1]); a.pre_z[p]=q[2]; a.pre_v[p]=*reinterpret_cast<const float4*>(rec+l.off_pre); }
  else if(a.pre_v){ a.pre_xy[p]=make_double2(d[0],d[1]); a.pre_z[p]=d[2]; a.pre_v[p]=*reinterpret_cast<const float4*>(rec+32); }
  if(l.off_m1)a.m1[p]=*reinterpret_cast<const float4*>(rec+l.off_m1);
  else if(a.m1)a.m1[p]=*reinterpret_cast<const float4*>(rec+32);
  if(l.off_tau){ const float2* s=reinterpret_cast<const float2*>(rec+l.off_tau); float2* t=reinterpret_cast<float2*>(a.tau+size_t(p)*6); t[0]=s[0]; t[1]=s[1]; t[2]=s[2]; }
  if(l.off_nor)a.nor[p]=*reinterpret_cast<const float4*>(rec+l.off_nor);
  if(B2S_CODE_TYPE(cod)<B2S_CODE_TYPE_FLOATING)atomicAdd(&ds->slab_nbound_tmp,1u);
}
__global__ void k_slab_unpack_commit(DivState* ds, uint32_t n){ ds->np+=n; ds->npb+=ds->slab_nbound_tmp; ds->slab_nbound_tmp=0; }

// keeps only the particles of this slab (initial distribution): everything outside [cxa,cxb) is marked OUTIGNORE
__global__ void __launch_bounds__(S_BS) k_slab_select(const DivState* __restrict__ ds, typecode* code, const uint32_t* __restrict__ dcell, uint32_t dcc, uint32_t cxa, uint32_t cxb){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=ds->np)return;
  const typecode c=code[p];
  if(B2S_CODE_SPECIAL(c)==B2S_CODE_PERIODIC){ code[p]=typecode(B2S_CODE_SETNORMAL(c)|B2S_CODE_OUTIGNORE); return; }   // periodic duplicates of the upload's divide
  if(B2S_CODE_SPECIAL(c)!=B2S_CODE_NORMAL)return;
  const uint32_t cx=B2S_DCEL_CELLX(dcc,dcell[p]);
  if(cx<cxa || cx>=cxb)code[p]=typecode(B2S_CODE_SETNORMAL(c)|B2S_CODE_OUTIGNORE);
}
__global__ void k_slab_fix_counts(DivState* ds){ ds->npf_out_total=0; }

static SlabArrays slab_arrays(b2s_ctx* ctx){
  SlabArrays a;
  a.idp=ctx->cur.idp; a.code=ctx->cur.code; a.dcell=ctx->cur.dcell; a.posxy=ctx->cur.posxy; a.posz=ctx->cur.posz; a.velrho=ctx->cur.velrho;
  a.pre_xy=(ctx->pre_active? ctx->pospre_xy: NULL); a.pre_z=(ctx->pre_active? ctx->pospre_z: NULL); a.pre_v=(ctx->pre_active? ctx->velrhopre: NULL);
  a.m1=ctx->velrhom1; a.tau=ctx->spstau; a.nor=ctx->boundnor;
  return a;
}

extern "C" int b2s_slab_config(b2s_ctx* ctx, int rank, int nranks, uint32_t cx_begin, uint32_t cx_end){
  if(!ctx || nranks<1 || rank<0 || rank>=nranks || cx_begin>=cx_end || cx_end>ctx->prm.domcells[0])return B2S_ERR_ARG;
  if(nranks>1 && cx_end-cx_begin<slab_halo_planes(ctx))return b2s_fail(ctx,B2S_ERR_ARG,"slab [%u,%u) is thinner than the halo (%u cell planes)",cx_begin,cx_end,slab_halo_planes(ctx));
  if(ctx->cte.periactive&1)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"periodic boundaries along x (the slab axis) are not supported with slab decomposition");
  if(ctx->cte.periactive && ctx->prm.tboundary==B2S_BC_MDBC && nranks>1)return b2s_fail(ctx,B2S_ERR_UNSUPPORTED,"periodic boundaries with mDBC are not supported with slab decomposition");
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  ctx->slab_on=true; ctx->slab_rank=rank; ctx->slab_nranks=nranks; ctx->slab_cxa=cx_begin; ctx->slab_cxb=cx_end;
  const size_t nb=b2s_blocks(ctx->cap,S_BS);
  if(!ctx->slab_counts){
    ctx->slab_counts_len=4*(nb+1);           // two sets (migrants, halos) for the merged exchange
    B2S_CUDA(ctx,cudaMalloc((void**)&ctx->slab_counts,ctx->slab_counts_len*4));
    B2S_CUDA(ctx,cudaMalloc((void**)&ctx->slab_flags,size_t(ctx->cap)*2));
  }
  // drop everything that does not belong to this slab and re-divide
  k_slab_select<<<b2s_blocks(ctx->np_host,S_BS),S_BS,0,ctx->stream>>>(ctx->div,ctx->cur.code,ctx->cur.dcell,ctx->cte.domcellcode,cx_begin,cx_end); B2S_LAUNCHED(ctx);
  ctx->bound_changed=true;
  int r=b2s_celldiv_impl(ctx); if(r)return r;
  k_slab_fix_counts<<<1,1,0,ctx->stream>>>(ctx->div); B2S_LAUNCHED(ctx);
  return b2s_sync_info(ctx);
}

// ---- re-cutting (dynamic load balance): histogram of owned particles per x cell plane, new slab range -------------------
__global__ void __launch_bounds__(S_BS) k_slab_hist(const DivState* __restrict__ ds, const typecode* __restrict__ code, const uint32_t* __restrict__ dcell,
  uint32_t dcc, uint32_t ncx, uint32_t* hist)
{
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=ds->np)return;
  if(B2S_CODE_SPECIAL(code[p])!=B2S_CODE_NORMAL)return;
  const uint32_t rc=dcell[p];
  if(rc==B2S_DCEL_MAPOUT)return;
  const uint32_t cx=B2S_DCEL_CELLX(dcc,rc);
  // particles of a warp mostly share a few planes: aggregate equal planes before the atomic
  const uint32_t peers=__match_any_sync(__activemask(),cx);
  // work units, not heads: a fluid particle weighs B2S_SLAB_WEIGHT_FLUID boundary particles (slab_cuts.h)
  const uint32_t w=(B2S_CODE_TYPE(code[p])<B2S_CODE_TYPE_FLOATING? B2S_SLAB_WEIGHT_BOUND: B2S_SLAB_WEIGHT_FLUID);
  uint32_t sum=0;
  for(uint32_t m=peers;m;m&=m-1)sum+=__shfl_sync(peers,w,__ffs(m)-1);
  if((__ffs(peers)-1)==int(threadIdx.x&31) && cx<ncx)atomicAdd(&hist[cx],sum);
}
extern "C" int b2s_slab_histogram(b2s_ctx* ctx, void* dev_u32_ncx){
  if(!ctx || !ctx->slab_on || !dev_u32_ncx)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  const uint32_t ncx=ctx->prm.domcells[0];
  B2S_CUDA(ctx,cudaMemsetAsync(dev_u32_ncx,0,size_t(ncx)*4,ctx->stream));
  if(ctx->np_host){
    k_slab_hist<<<b2s_blocks(ctx->np_host,S_BS),S_BS,0,ctx->stream>>>(ctx->div,ctx->cur.code,ctx->cur.dcell,ctx->cte.domcellcode,ncx,(uint32_t*)dev_u32_ncx); B2S_LAUNCHED(ctx);
  }
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
__global__ void __launch_bounds__(S_BS) k_slab_drop_ghosts(const DivState* __restrict__ ds, typecode* code){
  const uint32_t p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p<ds->np && B2S_CODE_SPECIAL(code[p])==B2S_CODE_PERIODIC)code[p]=typecode(B2S_CODE_SETNORMAL(code[p])|B2S_CODE_OUTIGNORE);
}
extern "C" int b2s_slab_drop_ghosts(b2s_ctx* ctx){
  if(!ctx || !ctx->slab_on)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  if(ctx->np_host){ k_slab_drop_ghosts<<<b2s_blocks(ctx->np_host,S_BS),S_BS,0,ctx->stream>>>(ctx->div,ctx->cur.code); B2S_LAUNCHED(ctx); }
  ctx->bound_changed=true;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
extern "C" int b2s_slab_set_range(b2s_ctx* ctx, uint32_t cx_begin, uint32_t cx_end){
  if(!ctx || !ctx->slab_on || cx_begin>=cx_end || cx_end>ctx->prm.domcells[0])return B2S_ERR_ARG;
  if(ctx->slab_nranks>1 && cx_end-cx_begin<slab_halo_planes(ctx))return b2s_fail(ctx,B2S_ERR_ARG,"slab [%u,%u) is thinner than the halo (%u cell planes)",cx_begin,cx_end,slab_halo_planes(ctx));
  ctx->slab_cxa=cx_begin; ctx->slab_cxb=cx_end;
  return B2S_OK;
}

extern "C" int b2s_slab_get_range(b2s_ctx* ctx, uint32_t* range2){
  if(!ctx || !ctx->slab_on || !range2)return B2S_ERR_ARG;
  range2[0]=ctx->slab_cxa; range2[1]=ctx->slab_cxb;
  return B2S_OK;
}

extern "C" int b2s_slab_record_bytes(b2s_ctx* ctx, int kind, uint32_t* bytes){
  if(!ctx || !bytes || kind<0 || kind>1)return B2S_ERR_ARG;
  *bytes=slab_layout(ctx,kind).bytes;
  return B2S_OK;
}

extern "C" int b2s_slab_pack(b2s_ctx* ctx, int kind, void* dev_left, void* dev_right, uint32_t cap_records, uint32_t* nleft, uint32_t* nright){
  if(!ctx || !ctx->slab_on || kind<0 || kind>1 || !nleft || !nright)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  cudaStream_t s=ctx->stream;
  const uint32_t n=ctx->np_host;
  *nleft=*nright=0;
  if(!n)return B2S_OK;
  const uint32_t nb=b2s_blocks(n,S_BS);
  const int hl=(ctx->slab_rank>0), hr=(ctx->slab_rank<ctx->slab_nranks-1);
  B2S_CUDA(ctx,cudaMemsetAsync(ctx->slab_counts,0,size_t(2)*(nb+1)*4,s));
  k_slab_flag<<<nb,S_BS,0,s>>>(ctx->div,ctx->cur.code,ctx->cur.dcell,ctx->cte.domcellcode,kind,hl,hr,ctx->slab_cxa,ctx->slab_cxb,slab_halo_planes(ctx),
    ctx->slab_flags,ctx->slab_counts,nb); B2S_LAUNCHED(ctx);
  b2s_launch_scan_excl(s,ctx->slab_counts,nb+1); B2S_LAUNCHED(ctx);
  b2s_launch_scan_excl(s,ctx->slab_counts+nb+1,nb+1); B2S_LAUNCHED(ctx);
  uint32_t tot[2]={0,0};
  B2S_CUDA(ctx,cudaMemcpyAsync(&tot[0],ctx->slab_counts+nb,4,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(&tot[1],ctx->slab_counts+2*nb+1,4,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaStreamSynchronize(s));
  if(tot[0]>cap_records || tot[1]>cap_records)return b2s_fail(ctx,B2S_ERR_NOMEM,"slab exchange buffer too small: need %u/%u records, have %u",tot[0],tot[1],cap_records);
  if((tot[0] && !dev_left) || (tot[1] && !dev_right))return B2S_ERR_ARG;
  const SlabLayout l=slab_layout(ctx,kind);
  k_slab_scatter<<<nb,S_BS,0,s>>>(ctx->div,ctx->slab_flags,ctx->slab_counts,nb,l,slab_arrays(ctx),kind,(char*)dev_left,(char*)dev_right,cap_records,typecode(B2S_CODE_OUTIGNORE)); B2S_LAUNCHED(ctx);
  B2S_CUDA(ctx,cudaGetLastError());
  *nleft=tot[0]; *nright=tot[1];
  return B2S_OK;
}

// Merged exchange: migrants AND halos of one interaction in one message per neighbour (one host sync, one count exchange, one
// payload exchange instead of two each).  A particle that migrates stays on the sender as a halo ghost of its new owner (it has
// just crossed the cut, so it lies inside the sender's halo region), hence the receiver need not send it back and the halo can be
// selected before the migrants arrive.  Needs slabs wider than the halo (a received migrant must not fall into the far halo band).
// Buffer per side: [counts[0|1] migrant records][counts[2|3] halo records].
extern "C" int b2s_slab_pack_both(b2s_ctx* ctx, void* dev_left, void* dev_right, uint64_t cap_bytes, uint32_t counts[4]){
  if(!ctx || !ctx->slab_on || !counts)return B2S_ERR_ARG;
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  cudaStream_t s=ctx->stream;
  counts[0]=counts[1]=counts[2]=counts[3]=0;
  const uint32_t n=ctx->np_host;
  if(!n)return B2S_OK;
  if(ctx->slab_nranks>1 && ctx->slab_cxb-ctx->slab_cxa<=slab_halo_planes(ctx))
    return b2s_fail(ctx,B2S_ERR_STATE,"merged exchange needs slabs wider than the halo (%u planes)",slab_halo_planes(ctx));
  const uint32_t nb=b2s_blocks(n,S_BS);
  const int hl=(ctx->slab_rank>0), hr=(ctx->slab_rank<ctx->slab_nranks-1);
  uint32_t* cM=ctx->slab_counts; uint32_t* cH=ctx->slab_counts+2*(nb+1);
  uint8_t* fM=ctx->slab_flags; uint8_t* fH=ctx->slab_flags+ctx->cap;
  B2S_CUDA(ctx,cudaMemsetAsync(ctx->slab_counts,0,size_t(4)*(nb+1)*4,s));
  k_slab_flag<<<nb,S_BS,0,s>>>(ctx->div,ctx->cur.code,ctx->cur.dcell,ctx->cte.domcellcode,1,hl,hr,ctx->slab_cxa,ctx->slab_cxb,slab_halo_planes(ctx),fM,cM,nb); B2S_LAUNCHED(ctx);
  k_slab_flag<<<nb,S_BS,0,s>>>(ctx->div,ctx->cur.code,ctx->cur.dcell,ctx->cte.domcellcode,0,hl,hr,ctx->slab_cxa,ctx->slab_cxb,slab_halo_planes(ctx),fH,cH,nb); B2S_LAUNCHED(ctx);
  b2s_launch_scan_excl(s,cM,nb+1); b2s_launch_scan_excl(s,cM+nb+1,nb+1); b2s_launch_scan_excl(s,cH,nb+1); b2s_launch_scan_excl(s,cH+nb+1,nb+1); ctx->launches+=4;
  uint32_t tot[4]={0,0,0,0};
  B2S_CUDA(ctx,cudaMemcpyAsync(&tot[0],cM+nb,4,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(&tot[1],cM+2*nb+1,4,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(&tot[2],cH+nb,4,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaMemcpyAsync(&tot[3],cH+2*nb+1,4,cudaMemcpyDeviceToHost,s));
  B2S_CUDA(ctx,cudaStreamSynchronize(s));
  const SlabLayout lM=slab_layout(ctx,1), lH=slab_layout(ctx,0);
  const uint64_t needl=uint64_t(tot[0])*lM.bytes+uint64_t(tot[2])*lH.bytes, needr=uint64_t(tot[1])*lM.bytes+uint64_t(tot[3])*lH.bytes;
  if(needl>cap_bytes || needr>cap_bytes)return b2s_fail(ctx,B2S_ERR_NOMEM,"slab exchange buffer too small: need %llu/%llu bytes, have %llu",(unsigned long long)needl,(unsigned long long)needr,(unsigned long long)cap_bytes);
  if((needl && !dev_left) || (needr && !dev_right))return B2S_ERR_ARG;
  if(tot[0]||tot[1]){ k_slab_scatter<<<nb,S_BS,0,s>>>(ctx->div,fM,cM,nb,lM,slab_arrays(ctx),1,(char*)dev_left,(char*)dev_right,0xffffffffu,typecode(B2S_CODE_PERIODIC)); B2S_LAUNCHED(ctx); ctx->bound_changed=true; }
  if(tot[2]||tot[3]){ k_slab_scatter<<<nb,S_BS,0,s>>>(ctx->div,fH,cH,nb,lH,slab_arrays(ctx),0,(char*)dev_left+size_t(tot[0])*lM.bytes,(char*)dev_right+size_t(tot[1])*lM.bytes,0xffffffffu,typecode(B2S_CODE_OUTIGNORE)); B2S_LAUNCHED(ctx); }
  B2S_CUDA(ctx,cudaGetLastError());
  for(int k=0;k<4;k++)counts[k]=tot[k];
  return B2S_OK;
}
extern "C" int b2s_slab_unpack(b2s_ctx* ctx, int kind, const void* dev_buf, uint32_t count);
extern "C" int b2s_slab_unpack_both(b2s_ctx* ctx, const void* dev_buf, uint32_t nmigrants, uint32_t nhalo){
  if(!ctx || !ctx->slab_on)return B2S_ERR_ARG;
  const uint32_t rbM=slab_layout(ctx,1).bytes;
  int r=b2s_slab_unpack(ctx,1,dev_buf,nmigrants); if(r)return r;
  return b2s_slab_unpack(ctx,0,(const char*)dev_buf+size_t(nmigrants)*rbM,nhalo);
}

extern "C" int b2s_slab_unpack(b2s_ctx* ctx, int kind, const void* dev_buf, uint32_t count){
  if(!ctx || !ctx->slab_on || kind<0 || kind>1)return B2S_ERR_ARG;
  if(!count)return B2S_OK;
  if(!dev_buf)return B2S_ERR_ARG;
  if(uint64_t(ctx->np_host)+count>ctx->cap)return b2s_fail(ctx,B2S_ERR_NOMEM,"particle capacity %u exceeded by %u incoming slab particles",ctx->cap,count);
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  const SlabLayout l=slab_layout(ctx,kind);
  k_slab_unpack<<<b2s_blocks(count,S_BS),S_BS,0,ctx->stream>>>(ctx->div,(const char*)dev_buf,count,l,slab_arrays(ctx),kind,ctx->cte); B2S_LAUNCHED(ctx);
  k_slab_unpack_commit<<<1,1,0,ctx->stream>>>(ctx->div,count); B2S_LAUNCHED(ctx);
  ctx->np_host+=count; ctx->npb_host+=count; ctx->npf_host+=count;   // host mirrors are upper bounds until the next sync
  ctx->bound_changed=true;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}

extern "C" int b2s_slab_dt_buffer(b2s_ctx* ctx, void** dev_u32x3){
  if(!ctx || !dev_u32x3)return B2S_ERR_ARG;
  *dev_u32x3=(void*)ctx->step;    // StepState starts with {velmax2_bits, acemax2_bits, viscdtmax_bits}: ordered uints, all-reduce with MAX
  return B2S_OK;
}

// ---- device-counted exchange: pack / unpack without a host synchronisation (driven by slab_comm.cu) -----------------------
// mode bit0: migrants, bit1: halos.  With both, a migrant stays on the sender as a halo ghost of its new owner (merged exchange,
// see b2s_slab_pack_both); migrants alone (re-cut rounds) are dropped by the sender's next divide.
int b2s_slab_pack_msg(b2s_ctx* ctx, char* dev_left, char* dev_right, uint32_t migcap, uint32_t halocap, int mode, uint32_t* dev_sent_total, int strict){
  cudaStream_t s=ctx->stream;
  const uint32_t n=ctx->np_host;
  const int hl=(ctx->slab_rank>0), hr=(ctx->slab_rank<ctx->slab_nranks-1);
  const uint32_t nb=b2s_blocks(std::max(n,1u),S_BS);
  if(4*(size_t(nb)+1)>ctx->slab_counts_len)return b2s_fail(ctx,B2S_ERR_STATE,"slab exchange: count buffer too small for %u particles",n);
  const bool mig=(mode&1)!=0, halo=(mode&2)!=0;
  if(halo && ctx->slab_nranks>1 && ctx->slab_cxb-ctx->slab_cxa<=slab_halo_planes(ctx))
    return b2s_fail(ctx,B2S_ERR_STATE,"the device-counted exchange needs slabs wider than the halo (%u planes)",slab_halo_planes(ctx));
  uint32_t* cM=ctx->slab_counts; uint32_t* cH=ctx->slab_counts+2*(nb+1);
  uint8_t* fM=ctx->slab_flags; uint8_t* fH=ctx->slab_flags+ctx->cap;
  B2S_CUDA(ctx,cudaMemsetAsync(ctx->slab_counts,0,size_t(4)*(nb+1)*4,s));
  const SlabLayout lM=slab_layout(ctx,1), lH=slab_layout(ctx,0);
  const SlabArrays a=slab_arrays(ctx);
  if(n){
    if(mig && halo){ k_slab_flag2<<<nb,S_BS,0,s>>>(ctx->div,ctx->cur.code,ctx->cur.dcell,ctx->cte.domcellcode,hl,hr,ctx->slab_cxa,ctx->slab_cxb,slab_halo_planes(ctx),fM,fH,cM,cH,nb); B2S_LAUNCHED(ctx); }
    else if(mig){ k_slab_flag<<<nb,S_BS,0,s>>>(ctx->div,ctx->cur.code,ctx->cur.dcell,ctx->cte.domcellcode,1,hl,hr,ctx->slab_cxa,ctx->slab_cxb,slab_halo_planes(ctx),fM,cM,nb); B2S_LAUNCHED(ctx); }
    else if(halo){ k_slab_flag<<<nb,S_BS,0,s>>>(ctx->div,ctx->cur.code,ctx->cur.dcell,ctx->cte.domcellcode,0,hl,hr,ctx->slab_cxa,ctx->slab_cxb,slab_halo_planes(ctx),fH,cH,nb); B2S_LAUNCHED(ctx); }
    // cM = [left | right] and cH = [left | right] are four consecutive segments of nb+1 counts
    if(mig && halo){ b2s_launch_scan_excl_seg(s,cM,nb+1,4); B2S_LAUNCHED(ctx); }
    else if(mig){ b2s_launch_scan_excl_seg(s,cM,nb+1,2); B2S_LAUNCHED(ctx); }
    else if(halo){ b2s_launch_scan_excl_seg(s,cH,nb+1,2); B2S_LAUNCHED(ctx); }
  }
  k_slab_header<<<1,32,0,s>>>(ctx->step,cM,cH,nb,migcap,halocap,hl? dev_left: NULL,hr? dev_right: NULL,dev_sent_total,strict); B2S_LAUNCHED(ctx);
  if(n && mig && halo){
    k_slab_scatter2<<<nb,S_BS,0,s>>>(ctx->div,fM,fH,cM,cH,nb,lM,lH,a,dev_left,dev_right,migcap,halocap,typecode(B2S_CODE_PERIODIC)); B2S_LAUNCHED(ctx);
  }
  else if(n){
    if(mig){ k_slab_scatter_dev<<<nb,S_BS,0,s>>>(ctx->div,fM,cM,nb,lM,a,1,dev_left,dev_right,migcap,NULL,NULL,0,0,typecode(halo? B2S_CODE_PERIODIC: B2S_CODE_OUTIGNORE)); B2S_LAUNCHED(ctx); }
    if(halo){ k_slab_scatter_dev<<<nb,S_BS,0,s>>>(ctx->div,fH,cH,nb,lH,a,0,dev_left,dev_right,halocap,cM+nb,cM+2*nb+1,migcap,lM.bytes,typecode(B2S_CODE_OUTIGNORE)); B2S_LAUNCHED(ctx); }
  }
  ctx->bound_changed=true;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
uint64_t b2s_slab_msg_bytes(b2s_ctx* ctx, uint32_t migcap, uint32_t halocap){
  return uint64_t(B2S_SLAB_HDR)+uint64_t(migcap)*slab_layout(ctx,1).bytes+uint64_t(halocap)*slab_layout(ctx,0).bytes;
}
// migrant records a message of msg_bytes can carry when it carries nothing else (re-cut rounds)
uint32_t b2s_slab_migrants_per_msg(b2s_ctx* ctx, uint64_t msg_bytes){ return uint32_t((msg_bytes-B2S_SLAB_HDR)/slab_layout(ctx,1).bytes); }
int b2s_slab_unpack_msg(b2s_ctx* ctx, const char* dev_msg, uint32_t migcap, uint32_t halocap){
  const uint32_t nmax=migcap+halocap;
  k_slab_unpack_msg<<<b2s_blocks(nmax,S_BS),S_BS,0,ctx->stream>>>(ctx->div,ctx->step,dev_msg,slab_layout(ctx,1),slab_layout(ctx,0),slab_arrays(ctx),ctx->cap,ctx->cte); B2S_LAUNCHED(ctx);
  k_slab_unpack_msg_commit<<<1,1,0,ctx->stream>>>(ctx->div,dev_msg,ctx->cap); B2S_LAUNCHED(ctx);
  // host mirrors: upper bounds until the next b2s_sync_info (never above the capacity: the device refuses what does not fit)
  ctx->np_host=uint32_t(std::min<uint64_t>(uint64_t(ctx->np_host)+nmax,ctx->cap));
  ctx->npb_host=uint32_t(std::min<uint64_t>(uint64_t(ctx->npb_host)+nmax,ctx->cap));
  ctx->npf_host=uint32_t(std::min<uint64_t>(uint64_t(ctx->npf_host)+nmax,ctx->cap));
  ctx->bound_changed=true;
  B2S_CUDA(ctx,cudaGetLastError());
  return B2S_OK;
}
