// slab_comm.cu -- the slab-decomposed time step in C++: NCCL send/recv over NVLink for migrants and halos, NCCL all-reduce for
// the dt maxima and the floating-body sums, no host synchronisation between the start of a step and its end.
//
// No reference counterpart: DualSPHysics v5.4 is single-GPU (AVAILABLE_MGPU is hard-false, src/source/DualSphDef.h:96-101); the step
// that is distributed here is JSphGpuSingle::ComputeStep_Sym / ComputeStep_Ver + the tail of the run loop
// (src/source/JSphGpuSingle.cpp:761-819,1003-1008).
//
// One process per GPU, one communicator per context (b2s_slab_comm_init; Python or the CLI only hand over the 128-byte NCCL id).
// Per interaction every rank
//   1. flags, scans and scatters its migrants and halo particles into ONE message per neighbour
//      [u32 nmigrants | u32 nhalo | pad][migrant records][halo records]                                 (slab.cu, k_slab_*_dev)
//   2. exchanges the messages with ncclSend/ncclRecv inside one group on the context's own stream.  Messages have a FIXED size
//      (the capacities of the buffers), the counts travel in the header and stay on the device, so neither side waits on the
//      host for the other's counts (round 1 paid two host round trips per exchange); NVLink moves the ~20 MB in ~25 us
//   3. appends what it received behind its own particles (the unpack kernel reads the counts from the header) and runs the
//      cell divide on owned + ghost particles.
// The host mirrors of the particle counts are upper bounds during the step (kernels guard with the device-side counts) and are
// refreshed by the ONE synchronisation at the end of the step, which also surfaces the error flags (capacity, dt).
// NCCL is resolved with dlopen at b2s_slab_comm_init: single-GPU runs neither need nor load it, and inside a Python process the
// library torch already loaded is reused.
// A second transport, "local", keeps every rank of a run in ONE process on ONE device (device copies, host-mediated reductions,
// a device synchronisation between phases): it exists so that the whole C++ step can be checked against the single-context run
// on a one-GPU box (tests/test_gpu_slab.py); it is never the product path.
#include "b2s_internal.h"
#include "slab_cuts.h"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <cstdlib>
#include <vector>

int b2s_slab_pack_msg(b2s_ctx* ctx, char* dev_left, char* dev_right, uint32_t migcap, uint32_t halocap, int mode, uint32_t* dev_sent_total, int strict);   // slab.cu
uint32_t b2s_slab_migrants_per_msg(b2s_ctx* ctx, uint64_t msg_bytes);
int b2s_slab_unpack_msg(b2s_ctx* ctx, const char* dev_msg, uint32_t migcap, uint32_t halocap);
uint64_t b2s_slab_msg_bytes(b2s_ctx* ctx, uint32_t migcap, uint32_t halocap);
int b2s_dtvariable_impl(b2s_ctx* ctx, int interstep);       // forces.cu
int b2s_slab_step_end_enqueue(b2s_ctx* ctx);                // ctx.cu

struct NcclApi {
  void* lib;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*,int,ncclUniqueId,int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void*,size_t,ncclDataType_t,int,ncclComm_t,cudaStream_t);
  ncclResult_t (*Recv)(void*,size_t,ncclDataType_t,int,ncclComm_t,cudaStream_t);
  ncclResult_t (*AllReduce)(const void*,void*,size_t,ncclDataType_t,ncclRedOp_t,ncclComm_t,cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl={NULL};
static const char* nccl_load(){
  if(g_nccl.lib)return NULL;
  void* h=dlopen("libnccl.so.2",RTLD_NOW|RTLD_GLOBAL);
  if(!h)h=dlopen("libnccl.so",RTLD_NOW|RTLD_GLOBAL);
  if(!h)return "libnccl.so.2 not found (needed for multi-GPU runs only)";
  NcclApi a; a.lib=h;
  #define B2S_NSYM(f,n) *(void**)(&a.f)=dlsym(h,n); if(!a.f)return "libnccl.so.2 lacks " n
  B2S_NSYM(GetUniqueId,"ncclGetUniqueId"); B2S_NSYM(CommInitRank,"ncclCommInitRank"); B2S_NSYM(CommDestroy,"ncclCommDestroy");
  B2S_NSYM(Send,"ncclSend"); B2S_NSYM(Recv,"ncclRecv"); B2S_NSYM(AllReduce,"ncclAllReduce");
  B2S_NSYM(GroupStart,"ncclGroupStart"); B2S_NSYM(GroupEnd,"ncclGroupEnd"); B2S_NSYM(GetErrorString,"ncclGetErrorString");
  #undef B2S_NSYM
  g_nccl=a;
  return NULL;
}
#define B2S_NCCL(ctx, expr) do{ ncclResult_t _e=(expr); if(_e!=ncclSuccess) return b2s_fail(ctx,B2S_ERR_STATE,"%s failed: %s",#expr,g_nccl.GetErrorString(_e)); }while(0)

struct SlabComm {
  ncclComm_t comm; bool nccl;
  std::vector<b2s_ctx*> group;          // local transport: every rank of the run, in rank order; with NCCL: this context alone
  char* send[2]; char* recv[2];         // [0] left neighbour, [1] right neighbour
  uint32_t migcap, halocap; uint64_t msg_bytes;
  uint32_t* sent_dev;                   // {migrants, halo particles} packed by the last exchange (device)
  uint32_t* hist_dev; uint32_t ncx;     // re-cut
  uint32_t recut_every; uint64_t nstep;
  std::vector<cudaEvent_t> ev; size_t evused;   // event pairs around the exchanges of the current step
  double ms_exchange; uint64_t nexchange, bytes_sent, nrecut;
  // B2S_SLAB_PROFILE=1: events at the phase boundaries of a step, ms per phase accumulated at the step's synchronisation
  bool profile; std::vector<cudaEvent_t> pev; std::vector<int> pid; size_t pused; double pms[16]; uint64_t psteps;
  // measured load of this rank since the last re-cut: ms in its own kernels (everything but the waits inside the collectives)
  bool timing; double busy_ms; uint64_t busy_steps; float* busy_dev;
};
enum { PH_FORCES=0, PH_REDUCE, PH_INTEGRATE, PH_PACK, PH_TRANSFER, PH_UNPACK, PH_DIVIDE, PH_STEPEND, PH_COUNT };
static const char* PH_NAME[PH_COUNT]={"forces(+mdbc,bound)","allreduce+dtvariable","integrate(+floating)","exchange:pack","exchange:send/recv","exchange:unpack","celldivide","step_end(damping,motion)"};
static inline SlabComm* comm_of(b2s_ctx* ctx){ return (SlabComm*)ctx->slab_comm; }

static int comm_alloc(b2s_ctx* ctx, uint32_t max_plane_particles){
  if(!ctx->slab_on)return b2s_fail(ctx,B2S_ERR_STATE,"b2s_slab_config first");
  if(ctx->slab_comm)return b2s_fail(ctx,B2S_ERR_STATE,"the context already has a slab transport");
  B2S_CUDA(ctx,cudaSetDevice(ctx->device));
  SlabComm* c=new SlabComm();
  c->comm=NULL; c->nccl=false; c->sent_dev=NULL; c->hist_dev=NULL; c->recut_every=0; c->nstep=0; c->evused=0;
  c->ms_exchange=0; c->nexchange=0; c->bytes_sent=0; c->nrecut=0;
  c->profile=(getenv("B2S_SLAB_PROFILE")!=NULL && getenv("B2S_SLAB_PROFILE")[0]=='1'); c->pused=0; c->psteps=0; for(int k=0;k<16;k++)c->pms[k]=0;
  c->timing=c->profile; c->busy_ms=0; c->busy_steps=0; c->busy_dev=NULL;
  for(int k=0;k<2;k++){ c->send[k]=NULL; c->recv[k]=NULL; }
  // halo: scelldiv (mDBC: 2*scelldiv) cell planes per side at the density of the fullest plane of the initial state, with half as
  // much again (WCSPH keeps the density within a few percent; what does not fit raises B2S_ERR_NOMEM at the step's sync);
  // migrants of one sub-step are a few thousand (a thirty-second of the halo capacity is ~4x what the collapsing column
  // sends at 10M particles per GPU)
  const uint32_t planes=uint32_t(ctx->prm.scelldiv)*(ctx->prm.tboundary==B2S_BC_MDBC? 2u: 1u);
  c->halocap=(3u*planes*std::max(max_plane_particles,1u))/2u+4096u;
  c->migcap=std::max(8192u,c->halocap/32u);
  c->msg_bytes=b2s_slab_msg_bytes(ctx,c->migcap,c->halocap);
  c->ncx=ctx->prm.domcells[0];
  for(int k=0;k<2;k++){
    const bool has=(k==0? ctx->slab_rank>0: ctx->slab_rank<ctx->slab_nranks-1);
    if(!has)continue;
    B2S_CUDA(ctx,cudaMalloc((void**)&c->send[k],c->msg_bytes)); B2S_CUDA(ctx,cudaMalloc((void**)&c->recv[k],c->msg_bytes));
    B2S_CUDA(ctx,cudaMemsetAsync(c->send[k],0,c->msg_bytes,ctx->stream)); B2S_CUDA(ctx,cudaMemsetAsync(c->recv[k],0,c->msg_bytes,ctx->stream));
  }
  B2S_CUDA(ctx,cudaMalloc((void**)&c->sent_dev,4*sizeof(uint32_t)));
  B2S_CUDA(ctx,cudaMemsetAsync(c->sent_dev,0,4*sizeof(uint32_t),ctx->stream));
  B2S_CUDA(ctx,cudaMalloc((void**)&c->hist_dev,size_t(c->ncx)*4));
  B2S_CUDA(ctx,cudaMalloc((void**)&c->busy_dev,size_t(2*ctx->slab_nranks)*4));
  ctx->slab_comm=c;
  return B2S_OK;
}

extern "C" int b2s_slab_comm_id(uint8_t* id128){
  if(!id128)return B2S_ERR_ARG;
  if(nccl_load())return B2S_ERR_STATE;
  ncclUniqueId id;
  if(g_nccl.GetUniqueId(&id)!=ncclSuccess)return B2S_ERR_STATE;
  static_assert(sizeof(ncclUniqueId)==128,"ncclUniqueId is 128 bytes");
  memcpy(id128,&id,128);
  return B2S_OK;
}
extern "C" int b2s_slab_comm_init(b2s_ctx* ctx, const uint8_t* id128, uint32_t max_plane_particles){
  if(!ctx || !id128)return B2S_ERR_ARG;
  if(const char* e=nccl_load())return b2s_fail(ctx,B2S_ERR_STATE,"%s",e);
  int r=comm_alloc(ctx,max_plane_particles); if(r)return r;
  SlabComm* c=comm_of(ctx);
  ncclUniqueId id; memcpy(&id,id128,128);
  B2S_NCCL(ctx,g_nccl.CommInitRank(&c->comm,ctx->slab_nranks,id,ctx->slab_rank));
  c->nccl=true; c->group.assign(1,ctx);
  return B2S_OK;
}
// Local transport (tests): all ranks of a run live in this process on one device.
extern "C" int b2s_slab_comm_init_local(b2s_ctx** ctxs, int n, uint32_t max_plane_particles){
  if(!ctxs || n<1)return B2S_ERR_ARG;
  for(int k=0;k<n;k++){
    if(!ctxs[k] || ctxs[k]->slab_rank!=k || ctxs[k]->slab_nranks!=n)return B2S_ERR_ARG;
    int r=comm_alloc(ctxs[k],max_plane_particles); if(r)return r;
  }
  for(int k=0;k<n;k++)comm_of(ctxs[k])->group.assign(ctxs,ctxs+n);
  return B2S_OK;
}
extern "C" int b2s_slab_comm_free(b2s_ctx* ctx){
  if(!ctx || !ctx->slab_comm)return B2S_OK;
  SlabComm* c=comm_of(ctx);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if(c->nccl && c->comm)g_nccl.CommDestroy(c->comm);
  for(int k=0;k<2;k++){ cudaFree(c->send[k]); cudaFree(c->recv[k]); }
  cudaFree(c->sent_dev); cudaFree(c->hist_dev); cudaFree(c->busy_dev);
  for(cudaEvent_t e:c->ev)cudaEventDestroy(e);
  for(cudaEvent_t e:c->pev)cudaEventDestroy(e);
  delete c; ctx->slab_comm=NULL;
  return B2S_OK;
}
extern "C" int b2s_slab_comm_set_recut(b2s_ctx* ctx, uint32_t every_steps){
  if(!ctx || !ctx->slab_comm)return B2S_ERR_ARG;
  comm_of(ctx)->recut_every=every_steps;
  if(every_steps)comm_of(ctx)->timing=true;      // the re-cut balances MEASURED time: events at the phase boundaries of every step
  return B2S_OK;
}
// out[0] ms spent in pack + transfer + unpack (CUDA events on the context's stream, so it includes waiting for a slower
// neighbour), [1] exchanges, [2] bytes sent, [3] re-cuts -- all since the last call with reset!=0; [4] migrants sent since
// b2s_slab_comm_init (device counter, read here with one synchronisation).
extern "C" int b2s_slab_comm_stats(b2s_ctx* ctx, double* out5, int reset){
  if(!ctx || !ctx->slab_comm || !out5)return B2S_ERR_ARG;
  SlabComm* c=comm_of(ctx);
  double* out4=out5;
  { uint32_t h[4]={0,0,0,0}; B2S_CUDA(ctx,cudaSetDevice(ctx->device)); B2S_CUDA(ctx,cudaMemcpyAsync(h,c->sent_dev,16,cudaMemcpyDeviceToHost,ctx->stream)); B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream)); out5[4]=double(h[2]); }
  out4[0]=c->ms_exchange; out4[1]=double(c->nexchange); out4[2]=double(c->bytes_sent); out4[3]=double(c->nrecut);
  if(reset){ c->ms_exchange=0; c->nexchange=0; c->bytes_sent=0; c->nrecut=0; c->psteps=0; for(int k=0;k<16;k++)c->pms[k]=0; }
  return B2S_OK;
}

// ---- phases: each is enqueued for every rank of the group; what needs another rank's data goes through xfer_* -------------
static int ev_mark(b2s_ctx* ctx){
  SlabComm* c=comm_of(ctx);
  if(c->evused==c->ev.size()){ cudaEvent_t e; B2S_CUDA(ctx,cudaEventCreate(&e)); c->ev.push_back(e); }
  B2S_CUDA(ctx,cudaEventRecord(c->ev[c->evused++],ctx->stream));
  return B2S_OK;
}
static int ev_collect(b2s_ctx* ctx){   // after a synchronisation: adds up the bracketed exchange times
  SlabComm* c=comm_of(ctx);
  for(size_t k=0;k+1<c->evused;k+=2){ float ms=0; if(cudaEventElapsedTime(&ms,c->ev[k],c->ev[k+1])==cudaSuccess)c->ms_exchange+=ms; }
  c->evused=0;
  return B2S_OK;
}

// profile: the time since the previous mark is booked to `phase`
static int ph_mark(b2s_ctx* ctx,int phase){
  SlabComm* c=comm_of(ctx);
  if(!c->timing)return B2S_OK;
  if(c->pused==c->pev.size()){ cudaEvent_t e; B2S_CUDA(ctx,cudaEventCreate(&e)); c->pev.push_back(e); c->pid.push_back(0); }
  c->pid[c->pused]=phase;
  B2S_CUDA(ctx,cudaEventRecord(c->pev[c->pused++],ctx->stream));
  return B2S_OK;
}
static void ph_collect(b2s_ctx* ctx){
  SlabComm* c=comm_of(ctx);
  if(!c->timing)return;
  for(size_t k=1;k<c->pused;k++){
    float ms=0;
    if(c->pid[k]>=0 && cudaEventElapsedTime(&ms,c->pev[k-1],c->pev[k])==cudaSuccess){
      c->pms[c->pid[k]]+=ms;
      if(c->pid[k]!=PH_REDUCE && c->pid[k]!=PH_TRANSFER)c->busy_ms+=ms;      // the waits for the other ranks sit in the collectives
    }
  }
  c->pused=0; c->psteps++; c->busy_steps++;
}
// ms per step and phase since b2s_slab_comm_init (B2S_SLAB_PROFILE=1), as text; empty without the profile
extern "C" int b2s_slab_comm_profile(b2s_ctx* ctx, char* out, uint32_t outlen){
  if(!ctx || !ctx->slab_comm || !out || !outlen)return B2S_ERR_ARG;
  SlabComm* c=comm_of(ctx);
  out[0]=0;
  if(!c->profile || !c->psteps)return B2S_OK;
  size_t w=0;
  for(int k=0;k<PH_COUNT && w+64<outlen;k++)w+=size_t(snprintf(out+w,outlen-w,"%s%s=%.4f",k? ", ": "",PH_NAME[k],c->pms[k]/double(c->psteps)));
  return B2S_OK;
}

static int group_sync(std::vector<b2s_ctx*>& g){ for(b2s_ctx* x:g){ B2S_CUDA(x,cudaSetDevice(x->device)); B2S_CUDA(x,cudaStreamSynchronize(x->stream)); } return B2S_OK; }

// messages to / from both neighbours
static int xfer_messages(std::vector<b2s_ctx*>& g){
  b2s_ctx* ctx=g[0]; SlabComm* c=comm_of(ctx);
  if(c->nccl){
    const int me=ctx->slab_rank, n=ctx->slab_nranks;
    B2S_NCCL(ctx,g_nccl.GroupStart());
    if(me>0){ B2S_NCCL(ctx,g_nccl.Send(c->send[0],c->msg_bytes,ncclChar,me-1,c->comm,ctx->stream)); B2S_NCCL(ctx,g_nccl.Recv(c->recv[0],c->msg_bytes,ncclChar,me-1,c->comm,ctx->stream)); }
    if(me<n-1){ B2S_NCCL(ctx,g_nccl.Send(c->send[1],c->msg_bytes,ncclChar,me+1,c->comm,ctx->stream)); B2S_NCCL(ctx,g_nccl.Recv(c->recv[1],c->msg_bytes,ncclChar,me+1,c->comm,ctx->stream)); }
    B2S_NCCL(ctx,g_nccl.GroupEnd());
    c->bytes_sent+=c->msg_bytes*uint64_t((me>0? 1: 0)+(me<n-1? 1: 0));
    return B2S_OK;
  }
  int r=group_sync(g); if(r)return r;
  for(size_t k=0;k<g.size();k++){
    SlabComm* ck=comm_of(g[k]);
    if(k>0){ B2S_CUDA(g[k],cudaMemcpy(ck->recv[0],comm_of(g[k-1])->send[1],ck->msg_bytes,cudaMemcpyDeviceToDevice)); ck->bytes_sent+=ck->msg_bytes; }
    if(k+1<g.size()){ B2S_CUDA(g[k],cudaMemcpy(ck->recv[1],comm_of(g[k+1])->send[0],ck->msg_bytes,cudaMemcpyDeviceToDevice)); ck->bytes_sent+=ck->msg_bytes; }
  }
  return B2S_OK;
}
// element-wise reduction of `count` 32-bit words at the same device address role on every rank (in place)
static int xfer_allreduce(std::vector<b2s_ctx*>& g, void* (*buf)(b2s_ctx*), size_t count, bool isfloat, bool ismax){
  b2s_ctx* ctx=g[0]; SlabComm* c=comm_of(ctx);
  if(!count)return B2S_OK;
  if(c->nccl){
    void* p=buf(ctx);
    B2S_NCCL(ctx,g_nccl.AllReduce(p,p,count,isfloat? ncclFloat32: ncclUint32,ismax? ncclMax: ncclSum,c->comm,ctx->stream));
    return B2S_OK;
  }
  int r=group_sync(g); if(r)return r;
  std::vector<uint32_t> acc(count,0u), tmp(count);
  std::vector<float> facc(count,0.f);
  bool first=true;
  for(b2s_ctx* x:g){
    B2S_CUDA(x,cudaMemcpy(tmp.data(),buf(x),count*4,cudaMemcpyDeviceToHost));
    for(size_t i=0;i<count;i++){
      if(isfloat){ float v; memcpy(&v,&tmp[i],4); facc[i]=(ismax? (first? v: std::max(facc[i],v)): facc[i]+v); }
      else acc[i]=(ismax? std::max(acc[i],tmp[i]): acc[i]+tmp[i]);
    }
    first=false;
  }
  for(b2s_ctx* x:g)B2S_CUDA(x,cudaMemcpy(buf(x),isfloat? (void*)facc.data(): (void*)acc.data(),count*4,cudaMemcpyHostToDevice));
  return B2S_OK;
}
static void* buf_dtmax(b2s_ctx* x){ return (void*)x->step; }          // StepState starts with the three maxima as ordered uints
static void* buf_ftsums(b2s_ctx* x){ return (void*)x->ft_allsums_dev; }
static void* buf_hist(b2s_ctx* x){ return (void*)comm_of(x)->hist_dev; }
static void* buf_sent(b2s_ctx* x){ return (void*)comm_of(x)->sent_dev; }
static void* buf_gsum(b2s_ctx* x){ return (void*)x->gauge_res_dev; }   // [results 4*ng | point masses] (gauges.cu gauges_upload)
static void* buf_gmax(b2s_ctx* x){ return (void*)(x->gauge_mass_dev+std::max<size_t>(x->gauge_points,1)); }   // raw maxima of the maxz gauges

#define FOR_GROUP(g,expr) do{ for(b2s_ctx* ctx:(g)){ B2S_CUDA(ctx,cudaSetDevice(ctx->device)); int _r=(expr); if(_r)return _r; } }while(0)

// RunCellDivide with the ghosts in place.  Periodic runs (y/z only with slabs) duplicate first, and that step sizes its work from
// exact host counts: they pay two synchronisations here (they are not the scaling target).
static int slab_divide(b2s_ctx* ctx){
  if(!ctx->cte.periactive)return b2s_celldiv_impl(ctx);
  int r=b2s_sync_info(ctx); if(r)return r;
  r=b2s_periodic_impl(ctx); if(r)return r;
  r=b2s_celldiv_impl(ctx); if(r)return r;
  return b2s_sync_info(ctx);
}

// migrants (+ halos) to the neighbours, then the cell divide with the ghosts in place
static int grp_exchange(std::vector<b2s_ctx*>& g, int mode, bool divide){
  // migrants alone (re-cut rounds) may fill the whole message, and what does not fit waits for the next round
  const bool recutround=(mode==1);
  const uint32_t migcap=(recutround? b2s_slab_migrants_per_msg(g[0],comm_of(g[0])->msg_bytes): comm_of(g[0])->migcap), halocap=(recutround? 0u: comm_of(g[0])->halocap);
  FOR_GROUP(g,ev_mark(ctx));
  FOR_GROUP(g,b2s_slab_pack_msg(ctx,comm_of(ctx)->send[0],comm_of(ctx)->send[1],migcap,halocap,mode,comm_of(ctx)->sent_dev,recutround? 0: 1));
  FOR_GROUP(g,ph_mark(ctx,PH_PACK));
  { int r=xfer_messages(g); if(r)return r; }
  FOR_GROUP(g,ph_mark(ctx,PH_TRANSFER));
  for(b2s_ctx* ctx:g){
    B2S_CUDA(ctx,cudaSetDevice(ctx->device));
    SlabComm* c=comm_of(ctx);
    for(int k=0;k<2;k++)if(c->recv[k]){ int r=b2s_slab_unpack_msg(ctx,c->recv[k],migcap,halocap); if(r)return r; }
    c->nexchange++;
  }
  FOR_GROUP(g,ev_mark(ctx));
  FOR_GROUP(g,ph_mark(ctx,PH_UNPACK));
  if(divide){ FOR_GROUP(g,slab_divide(ctx)); FOR_GROUP(g,ph_mark(ctx,PH_DIVIDE)); }
  return B2S_OK;
}
static int grp_interaction(std::vector<b2s_ctx*>& g, int interstep){
  FOR_GROUP(g,b2s_slab_forces(ctx,interstep));
  FOR_GROUP(g,ph_mark(ctx,PH_FORCES));
  { int r=xfer_allreduce(g,buf_dtmax,3,false,true); if(r)return r; }
  FOR_GROUP(g,b2s_dtvariable_impl(ctx,interstep));
  FOR_GROUP(g,ph_mark(ctx,PH_REDUCE));
  return B2S_OK;
}
static int grp_floating(std::vector<b2s_ctx*>& g, bool predictor){
  if(!g[0]->ftcount)return B2S_OK;
  for(b2s_ctx* ctx:g){ B2S_CUDA(ctx,cudaSetDevice(ctx->device)); void* p; uint32_t n; int r=b2s_slab_ft_sums(ctx,&p,&n); if(r)return r; }
  { int r=xfer_allreduce(g,buf_ftsums,6u*g[0]->ftcount,true,false); if(r)return r; }
  FOR_GROUP(g,b2s_slab_ft_apply(ctx,predictor? 1: 0));
  return B2S_OK;
}

// Gauges of a slab run (gauges.cu b2s_gauges_run): evaluated after the last exchange and divide of the step, when every rank holds
// fresh ghosts.  A measurement point belongs to the rank whose slab holds its cell, a boundary particle of a force gauge to its owner;
// what the ranks found is summed (exactly one of them contributes to a point), the maxz gauges take the maximum; every rank ends
// with the same results.  (One context evaluates before the end-of-step divide like the reference, JSphCpuSingle.cpp:1192: same
// particles, the sums in another order.)
static int grp_gauges(std::vector<b2s_ctx*>& g){
  const size_t ng=g[0]->gauge_types.size();
  if(!ng)return B2S_OK;
  FOR_GROUP(g,b2s_gauges_run(ctx,1));
  { int r=xfer_allreduce(g,buf_gsum,4*ng+std::max<size_t>(g[0]->gauge_points,1),true,false); if(r)return r; }
  { int r=xfer_allreduce(g,buf_gmax,ng,true,true); if(r)return r; }
  FOR_GROUP(g,b2s_gauges_run(ctx,2));
  return B2S_OK;
}

// Dynamic load balance: new cut planes from the all-reduced work histogram along x, then the particles outside their slab's new
// range travel (one neighbour hop per round) until nobody has migrants left; halos and divide at the end.
static int grp_recut(std::vector<b2s_ctx*>& g){
  const int n=g[0]->slab_nranks;
  FOR_GROUP(g,b2s_slab_drop_ghosts(ctx));
  FOR_GROUP(g,slab_divide(ctx));
  FOR_GROUP(g,b2s_slab_histogram(ctx,comm_of(ctx)->hist_dev));
  const uint32_t ncx=comm_of(g[0])->ncx;
  { int r=xfer_allreduce(g,buf_hist,ncx,false,false); if(r)return r; }
  std::vector<uint32_t> h(ncx);
  { b2s_ctx* ctx=g[0]; B2S_CUDA(ctx,cudaSetDevice(ctx->device)); B2S_CUDA(ctx,cudaMemcpyAsync(h.data(),comm_of(ctx)->hist_dev,size_t(ncx)*4,cudaMemcpyDeviceToHost,ctx->stream)); B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream)); }
  std::vector<uint64_t> work(h.begin(),h.end());
  { // Feedback from the clock: every rank knows how long its own kernels took per step since the last re-cut.  The work units of
    // its planes are priced at (its ms) / (its units), so the new cuts balance predicted TIME; the static weights (slab_cuts.h)
    // only shape the distribution inside a slab.  (A rank full of free-surface fluid is cheaper per particle than one deep inside
    // the column: measured 1.90 against 1.99 us per fluid particle and interaction at 10M particles per GPU.)
    std::vector<float> b(size_t(2*n),0.f);
    bool have=true;
    for(b2s_ctx* x:g){
      SlabComm* cx=comm_of(x);
      std::vector<float> mine(size_t(2*n),0.f);
      mine[x->slab_rank]=(cx->busy_steps? float(cx->busy_ms/double(cx->busy_steps)): 0.f);
      mine[n+x->slab_rank]=float(x->slab_cxa);
      B2S_CUDA(x,cudaSetDevice(x->device));
      B2S_CUDA(x,cudaMemcpyAsync(cx->busy_dev,mine.data(),size_t(2*n)*4,cudaMemcpyHostToDevice,x->stream));
      B2S_CUDA(x,cudaStreamSynchronize(x->stream));      // `mine` leaves scope; a re-cut synchronises anyway
    }
    { int r=xfer_allreduce(g,[](b2s_ctx* x)->void*{ return (void*)comm_of(x)->busy_dev; },size_t(2*n),true,false); if(r)return r; }
    { b2s_ctx* ctx=g[0]; B2S_CUDA(ctx,cudaSetDevice(ctx->device)); B2S_CUDA(ctx,cudaMemcpyAsync(b.data(),comm_of(ctx)->busy_dev,size_t(2*n)*4,cudaMemcpyDeviceToHost,ctx->stream)); B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream)); }
    for(int r=0;r<n;r++)if(!(b[r]>0.f))have=false;
    if(have){
      for(int r=0;r<n;r++){
        const uint32_t xa=uint32_t(b[n+r]), xb=(r+1<n? uint32_t(b[n+r+1]): ncx);
        uint64_t units=0; for(uint32_t x=xa;x<xb && x<ncx;x++)units+=h[x];
        const double price=(units? double(b[r])/double(units): 0.0);           // ms per work unit on rank r
        for(uint32_t x=xa;x<xb && x<ncx;x++)work[x]=uint64_t(double(h[x])*price*1e9+0.5);
      }
    }
    for(b2s_ctx* x:g){ comm_of(x)->busy_ms=0; comm_of(x)->busy_steps=0; }
  }
  const uint32_t planes=uint32_t(g[0]->prm.scelldiv)*(g[0]->prm.tboundary==B2S_BC_MDBC? 2u: 1u);
  // the device-counted exchange needs slabs WIDER than the halo
  const std::vector<uint32_t> cuts=b2s_slab_compute_cuts(work,n,planes+1u);
  for(b2s_ctx* ctx:g){ int r=b2s_slab_set_range(ctx,cuts[ctx->slab_rank],cuts[ctx->slab_rank+1]); if(r)return r; }
  // a round moves what one message holds one neighbour hop further; no divide in between (arrivals are appended with their cell
  // and a normal code, so the next round's flags see them; what left is marked and skipped)
  for(int round=0;;round++){
    if(round>64+n)return b2s_fail(g[0],B2S_ERR_STATE,"re-cut: migrants left after %d rounds",round);
    { int r=grp_exchange(g,1,false); if(r)return r; }
    { int r=xfer_allreduce(g,buf_sent,2,false,false); if(r)return r; }
    uint32_t moved[2]={0,0};
    { b2s_ctx* ctx=g[0]; B2S_CUDA(ctx,cudaSetDevice(ctx->device)); B2S_CUDA(ctx,cudaMemcpyAsync(moved,comm_of(ctx)->sent_dev,8,cudaMemcpyDeviceToHost,ctx->stream)); B2S_CUDA(ctx,cudaStreamSynchronize(ctx->stream)); }
    if(!moved[0])break;
  }
  for(b2s_ctx* ctx:g)comm_of(ctx)->nrecut++;
  return grp_exchange(g,2,true);
}

static int grp_step(std::vector<b2s_ctx*>& g, double* dt_out){
  const bool sym=(g[0]->prm.tstep==B2S_STEP_SYMPLECTIC);
  int r;
  FOR_GROUP(g,ph_mark(ctx,-1));
  if(sym){
    r=grp_interaction(g,B2S_INTER_SYMPREDICTOR); if(r)return r;
    FOR_GROUP(g,b2s_slab_integrate(ctx,1));
    r=grp_floating(g,true); if(r)return r;
    FOR_GROUP(g,ph_mark(ctx,PH_INTEGRATE));
    r=grp_exchange(g,3,true); if(r)return r;
    r=grp_interaction(g,B2S_INTER_SYMCORRECTOR); if(r)return r;
    FOR_GROUP(g,b2s_slab_integrate(ctx,2));
    r=grp_floating(g,false); if(r)return r;
    FOR_GROUP(g,ph_mark(ctx,PH_INTEGRATE));
  }
  else{
    r=grp_interaction(g,B2S_INTER_VERLET); if(r)return r;
    FOR_GROUP(g,b2s_slab_integrate(ctx,0));
    r=grp_floating(g,false); if(r)return r;
    FOR_GROUP(g,ph_mark(ctx,PH_INTEGRATE));
  }
  FOR_GROUP(g,b2s_slab_step_end_enqueue(ctx));
  FOR_GROUP(g,ph_mark(ctx,PH_STEPEND));
  r=grp_exchange(g,3,true); if(r)return r;
  r=grp_gauges(g); if(r)return r;
  SlabComm* c0=comm_of(g[0]);
  for(b2s_ctx* ctx:g)comm_of(ctx)->nstep++;
  if(c0->recut_every && c0->nstep%c0->recut_every==0){ r=grp_recut(g); if(r)return r; }
  // the one synchronisation of the step: exact counts for the next launch bounds, error flags of every kernel of the step
  FOR_GROUP(g,b2s_sync_info(ctx));
  FOR_GROUP(g,ev_collect(ctx));
  for(b2s_ctx* x:g)ph_collect(x);
  if(dt_out)*dt_out=g[0]->step_h->dt;
  return B2S_OK;
}

static int group_of(b2s_ctx* ctx, std::vector<b2s_ctx*>*& g){
  if(!ctx)return B2S_ERR_ARG;
  if(!ctx->slab_comm)return b2s_fail(ctx,B2S_ERR_STATE,"no slab transport (b2s_slab_comm_init)");
  g=&comm_of(ctx)->group;
  return B2S_OK;
}
// With NCCL every rank calls these on its own context (collectively); with the local transport one call on any context of the
// group drives all of them.
extern "C" int b2s_slab_exchange(b2s_ctx* ctx){
  std::vector<b2s_ctx*>* g; int r=group_of(ctx,g); if(r)return r;
  r=grp_exchange(*g,3,true); if(r)return r;
  for(b2s_ctx* x:*g){ r=b2s_sync_info(x); if(r)return r; ev_collect(x); }
  return B2S_OK;
}
extern "C" int b2s_slab_gauges_compute(b2s_ctx* ctx){
  std::vector<b2s_ctx*>* g; int r=group_of(ctx,g); if(r)return r;
  r=grp_gauges(*g); if(r)return r;
  for(b2s_ctx* x:*g){ r=b2s_sync_info(x); if(r)return r; }
  return B2S_OK;
}
extern "C" int b2s_slab_step(b2s_ctx* ctx, double* dt_out){
  std::vector<b2s_ctx*>* g; int r=group_of(ctx,g); if(r)return r;
  return grp_step(*g,dt_out);
}
extern "C" int b2s_slab_run_steps(b2s_ctx* ctx, uint32_t nsteps, double* time_out){
  std::vector<b2s_ctx*>* g; int r=group_of(ctx,g); if(r)return r;
  for(uint32_t k=0;k<nsteps;k++){ r=grp_step(*g,NULL); if(r)return r; }
  if(time_out)*time_out=ctx->step_h->timestep;
  return B2S_OK;
}
extern "C" int b2s_slab_recut(b2s_ctx* ctx, uint32_t* cuts_out, uint32_t ncuts){
  std::vector<b2s_ctx*>* g; int r=group_of(ctx,g); if(r)return r;
  r=grp_recut(*g); if(r)return r;
  for(b2s_ctx* x:*g){ r=b2s_sync_info(x); if(r)return r; ev_collect(x); }
  if(cuts_out && ncuts==2){ cuts_out[0]=ctx->slab_cxa; cuts_out[1]=ctx->slab_cxb; }
  return B2S_OK;
}
