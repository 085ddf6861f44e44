// frozen_device.cpp -- TEST INFRASTRUCTURE, never shipped: a stand-in for the device half of the C-ABI (include/b2s.h) that keeps the
// particles exactly as they were uploaded and only advances the clock.  Linked with the real host driver (csrc/host/*.cpp,
// csrc/config.cpp) it lets the CPU suite run everything the driver does around the device -- case parsing, the run loop with its
// PART schedule, every output file, the restart -- without a GPU (tests/test_host_driver_frozen.py).  The product library has no
// such path: without a CUDA device b2s_create fails.
#include "b2s.h"
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

struct b2s_ctx {
  b2s_params prm; uint32_t np=0; double time=0, dtpre=0; uint32_t nstep=0; int ngauges=0;
  std::vector<uint32_t> idp; std::vector<uint16_t> code; std::vector<double> pos; std::vector<float> velrho;
  std::vector<float> nor4;            // boundary normals as the device keeps them (float4 per particle, zero for the fluid)
  std::vector<b2s_floating> fts;
};
static const double FROZEN_DT=1.0/1024.0;   // exact in binary: the PART schedule of a test is the same on every machine

extern "C" {
int b2s_create(const b2s_params* p,int,uint32_t,b2s_ctx** out){ if(!p||!out)return B2S_ERR_ARG; b2s_ctx* c=new b2s_ctx; c->prm=*p; *out=c; return B2S_OK; }
int b2s_destroy(b2s_ctx* c){ delete c; return B2S_OK; }
const char* b2s_last_error(const b2s_ctx*){ return "frozen device"; }
int b2s_upload_particles(b2s_ctx* c,uint32_t np,const uint32_t* idp,const uint16_t* code,const double* pos,const float* velrho,const float* boundnor){
  c->np=np; c->idp.assign(idp,idp+np); c->code.assign(code,code+np); c->pos.assign(pos,pos+size_t(np)*3); c->velrho.assign(velrho,velrho+size_t(np)*4);
  c->nor4.assign(size_t(np)*4,0.f);
  if(boundnor){   // one float3 per boundary particle, the boundary comes first
    const uint32_t nbound=c->prm.casenp-c->prm.casenfluid;
    uint32_t nb=0; while(nb<np && idp[nb]<nbound)nb++;
    for(uint32_t i=0;i<nb;i++)for(int k=0;k<3;k++)c->nor4[size_t(i)*4+k]=boundnor[size_t(i)*3+k];
  }
  return B2S_OK;
}
int b2s_get_array(b2s_ctx* c,const char* name,void* dst,size_t capacity_bytes,size_t* count_out){
  const void* src=NULL; size_t bytes=0;
  if(!strcmp(name,"idp")){ src=c->idp.data(); bytes=size_t(c->np)*4; }
  else if(!strcmp(name,"boundnor") && c->prm.tboundary==B2S_BC_MDBC){ src=c->nor4.data(); bytes=size_t(c->np)*16; }
  else return B2S_ERR_ARG;
  if(capacity_bytes<bytes)return B2S_ERR_ARG;
  memcpy(dst,src,bytes); if(count_out)*count_out=c->np;
  return B2S_OK;
}
int b2s_download_particles_byid(b2s_ctx* c,uint32_t* np_out,uint32_t* idp,double* pos,float* vel,float* rhop){
  std::vector<uint32_t> o(c->np); std::iota(o.begin(),o.end(),0u);
  std::sort(o.begin(),o.end(),[&](uint32_t a,uint32_t b){ return c->idp[a]<c->idp[b]; });
  for(uint32_t k=0;k<c->np;k++){
    const uint32_t i=o[k]; idp[k]=c->idp[i];
    for(int d=0;d<3;d++){ pos[size_t(k)*3+d]=c->pos[size_t(i)*3+d]; vel[size_t(k)*3+d]=c->velrho[size_t(i)*4+d]; }
    rhop[k]=c->velrho[size_t(i)*4+3];
  }
  if(np_out)*np_out=c->np;
  return B2S_OK;
}
int b2s_get_pos_byid(b2s_ctx* c,uint32_t n,const uint32_t* ids,double* pos_out){
  for(uint32_t k=0;k<n;k++){
    const auto it=std::find(c->idp.begin(),c->idp.end(),ids[k]);
    if(it==c->idp.end())return B2S_ERR_ARG;
    const size_t i=size_t(it-c->idp.begin());
    for(int d=0;d<3;d++)pos_out[size_t(k)*3+d]=c->pos[i*3+d];
  }
  return B2S_OK;
}
int b2s_get_info(b2s_ctx* c,b2s_info* out){
  memset(out,0,sizeof(*out));
  const uint32_t npb=c->prm.casenfixed+c->prm.casenmoving;
  out->np=c->np; out->npb=npb; out->npbok=npb; out->nstep=c->nstep; out->timestep=c->time; out->dt_last=FROZEN_DT;
  out->symplectic_dtpre=(c->nstep? FROZEN_DT: c->dtpre);
  out->ncx=c->prm.map_cells[0]; out->ncy=c->prm.map_cells[1]; out->ncz=c->prm.map_cells[2]; out->nct=out->ncx*out->ncy*out->ncz;
  return B2S_OK;
}
int b2s_set_time(b2s_ctx* c,double t,double dtpre){ c->time=t; c->dtpre=dtpre; return B2S_OK; }
int b2s_compute_step(b2s_ctx* c,double* dt_out){ c->time+=FROZEN_DT; c->nstep++; if(dt_out)*dt_out=FROZEN_DT; return B2S_OK; }
int b2s_set_interaction_variant(b2s_ctx*,int){ return B2S_OK; }
int b2s_debug_set_tile_cap(b2s_ctx*,uint32_t){ return B2S_OK; }
int b2s_slab_config(b2s_ctx*,int,int,uint32_t,uint32_t){ return B2S_OK; }
int b2s_slab_comm_id(uint8_t* id){ memset(id,0,128); return B2S_OK; }
int b2s_slab_comm_init(b2s_ctx*,const uint8_t*,uint32_t){ return B2S_OK; }
int b2s_slab_exchange(b2s_ctx*){ return B2S_OK; }
int b2s_slab_step(b2s_ctx* c,double* dt_out){ return b2s_compute_step(c,dt_out); }
int b2s_slab_gauges_compute(b2s_ctx*){ return B2S_OK; }
int b2s_add_motion(b2s_ctx*,const b2s_motion*){ return B2S_OK; }
int b2s_out_enable(b2s_ctx*,uint32_t){ return B2S_OK; }
int b2s_out_fetch(b2s_ctx*,uint32_t,uint32_t* count_out,uint32_t*,double*,float*,float*,uint8_t*){ if(count_out)*count_out=0; return B2S_OK; }
int b2s_add_damping_plane(b2s_ctx*,const double*,const double*,float,float,const float*,const double*){ return B2S_OK; }
int b2s_add_damping_box(b2s_ctx*,const double*,const double*,const double*,const double*,float,float,const float*,uint32_t){ return B2S_OK; }
int b2s_set_floatings(b2s_ctx* c,uint32_t n,const b2s_floating* f){ c->fts.assign(f,f+n); return B2S_OK; }
int b2s_get_floatings(b2s_ctx* c,uint32_t n,b2s_floating* f){ if(n!=c->fts.size())return B2S_ERR_ARG; std::copy(c->fts.begin(),c->fts.end(),f); return B2S_OK; }
int b2s_gauge_add_velocity(b2s_ctx* c,const double*,int* id){ if(id)*id=c->ngauges; c->ngauges++; return B2S_OK; }
int b2s_gauge_add_swl(b2s_ctx* c,const double*,const double*,double,float,int* id){ if(id)*id=c->ngauges; c->ngauges++; return B2S_OK; }
int b2s_gauge_add_force(b2s_ctx* c,uint32_t,uint32_t,uint32_t,int* id){ if(id)*id=c->ngauges; c->ngauges++; return B2S_OK; }
int b2s_gauge_add_maxz(b2s_ctx* c,const double*,double,float,int* id){ if(id)*id=c->ngauges; c->ngauges++; return B2S_OK; }
int b2s_gauges_compute(b2s_ctx*){ return B2S_OK; }
int b2s_gauge_result(b2s_ctx* c,int id,float out[3]){ if(id<0||id>=c->ngauges)return B2S_ERR_ARG; out[0]=out[1]=out[2]=0.f; return B2S_OK; }
}
