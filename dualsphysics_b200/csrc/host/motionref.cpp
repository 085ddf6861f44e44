// motionref.cpp -- see motionref.h.
#include "motionref.h"
#include "bi4.h"
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <stdexcept>
namespace b2s {

namespace {
struct P3 { double x,y,z; };
inline double dist(const P3& a,const P3& b){ const double dx=a.x-b.x,dy=a.y-b.y,dz=a.z-b.z; return std::sqrt(dx*dx+dy*dy+dz*dz); }
// Distance of `pt` to the line (a,b) as twice the triangle area over the base (FunGeo3d.h:494-500, FunGeo3d.cpp:245-262).
inline double line_point_dist(const P3& pt,const P3& a,const P3& b){
  const double qx=a.x-pt.x,qy=a.y-pt.y,qz=a.z-pt.z, rx=b.x-pt.x,ry=b.y-pt.y,rz=b.z-pt.z;
  const double vi=qy*rz-ry*qz, vj=rx*qz-qx*rz, vk=qx*ry-rx*qy;
  return (0.5*std::sqrt(vi*vi+vj*vj+vk*vk)*2)/dist(a,b);
}
// Index in the ascending idp[] of particle `id`, or n when it is absent (excluded floating particle).
inline uint32_t find_id(uint32_t n,const uint32_t* idp,uint32_t id){
  const uint32_t* p=std::lower_bound(idp,idp+n,id);
  return (p!=idp+n && *p==id)? uint32_t(p-idp): n;
}
}

void MotionRefSave::add_body(bool floating,uint16_t mkbound,uint32_t begin,uint32_t np){
  if(!floating && nfloat)throw std::runtime_error("PartMotionRef: moving block after floating blocks");
  MotRefBody b; b.mkbound=mkbound; b.begin=begin; b.np=np; b.nid=0;
  for(int c=0;c<3;c++){ b.id[c]=UINT_MAX; b.ps[c][0]=b.ps[c][1]=b.ps[c][2]=DBL_MAX; b.dis[c]=DBL_MAX; }   // JComputeMotionRef::AddMkBlock
  bodies.push_back(b); if(floating)nfloat++; else nmoving++;
}

// JComputeMotionRef::ComputeRefPoints (JComputeMotionRef.cpp:127-236): the particles nearest to the 8 corners of the body's
// bounding box (first one wins on ties, repeated ones dropped); then [1] = the one furthest from [0]; then [2] = the one that
// maximises the distance of [0] to the line ([1],candidate) -- the argument order of the reference's call, kept as it is.
void MotionRefSave::config(uint32_t n,const uint32_t* idp,const double* pos){
  for(MotRefBody& b:bodies){
    auto P=[&](uint32_t k){ return P3{pos[size_t(k)*3],pos[size_t(k)*3+1],pos[size_t(k)*3+2]}; };
    // ids ascending: the particles of the body that are still in the simulation are the slots [k0,k1)
    const uint32_t k0=uint32_t(std::lower_bound(idp,idp+n,b.begin)-idp), k1=uint32_t(std::lower_bound(idp,idp+n,b.begin+b.np)-idp);
    P3 mn{DBL_MAX,DBL_MAX,DBL_MAX},mx{-DBL_MAX,-DBL_MAX,-DBL_MAX};
    for(uint32_t k=k0;k<k1;k++){
      const P3 p=P(k);
      mn.x=std::min(mn.x,p.x); mn.y=std::min(mn.y,p.y); mn.z=std::min(mn.z,p.z);
      mx.x=std::max(mx.x,p.x); mx.y=std::max(mx.y,p.y); mx.z=std::max(mx.z,p.z);
    }
    P3 plim[8]; uint32_t idlim[8]; unsigned nlim=0;
    for(unsigned c=0;c<8;c++){
      const bool xmax=(c==1||c==2||c==5||c==6), ymax=(c==2||c==3||c==6||c==7), zmax=(c>=4);
      const P3 corner{xmax? mx.x: mn.x, ymax? mx.y: mn.y, zmax? mx.z: mn.z};
      uint32_t best=b.begin; P3 pbest{0,0,0}; double dbest=DBL_MAX;
      for(uint32_t k=k0;k<k1;k++){
        const P3 p=P(k); const double d=dist(p,corner);
        if(dbest>d){ dbest=d; best=idp[k]; pbest=p; }
      }
      bool rep=false; for(unsigned r=0;r<nlim && !rep;r++)rep=(idlim[r]==best);
      if(!rep){ idlim[nlim]=best; plim[nlim]=pbest; nlim++; }
    }
    if(nlim>1){
      unsigned sel=1; double d=dist(plim[0],plim[1]);
      for(unsigned c=2;c<nlim;c++){ const double d2=dist(plim[0],plim[c]); if(d<d2){ d=d2; sel=c; } }
      std::swap(idlim[1],idlim[sel]); std::swap(plim[1],plim[sel]);
    }
    if(nlim>2){
      unsigned sel=2; double d=line_point_dist(plim[0],plim[1],plim[2]);
      for(unsigned c=3;c<nlim;c++){ const double d2=line_point_dist(plim[0],plim[1],plim[c]); if(d<d2){ d=d2; sel=c; } }
      std::swap(idlim[2],idlim[sel]); std::swap(plim[2],plim[sel]);
    }
    b.nid=(nlim>3? 3: nlim);
    for(unsigned c=0;c<b.nid;c++){
      b.id[c]=idlim[c]; b.ps[c][0]=plim[c].x; b.ps[c][1]=plim[c].y; b.ps[c][2]=plim[c].z;
      b.dis[c]=dist(plim[0],plim[c])*1.1;
    }
  }
}

namespace {
// Head of PartMotionRef.ibi4 / PartMotionRef2.ibi4 (JPartMotRefBi4Save::Config, JPartMotRefBi4Save.cpp:149-191).  The arrays live in `st`.
struct HeadStore { std::vector<uint16_t> mkb; std::vector<uint32_t> nid,id; std::vector<double> ps,dis; };
void make_head(Bi4Item& hd,HeadStore& st,const std::vector<MotRefBody>& bodies,unsigned nmoving,unsigned nfloat,
               const std::string& appname,bool mainfile,double timeout,unsigned mkboundfirst){
  const unsigned nmk=unsigned(bodies.size());
  hd.set_text("AppName",appname); hd.set_uint("FormatVer",230729); hd.set_bool("MainFile",mainfile); hd.set_double("TimeOut",timeout);
  { Bi4Value v; v.name="MkBoundFirst"; v.type=BT_USHORT; v.i=0; v.u=mkboundfirst; v.d=0; v.d3[0]=v.d3[1]=v.d3[2]=0; hd.values.push_back(v); }
  hd.set_uint("MkMovingCount",nmoving); hd.set_uint("MkFloatCount",nfloat); hd.set_uint("MkCount",nmk);
  st.mkb.resize(nmk); st.nid.resize(nmk); st.id.resize(size_t(nmk)*3); st.ps.resize(size_t(nmk)*9); st.dis.resize(size_t(nmk)*3);
  for(unsigned c=0;c<nmk;c++){
    const MotRefBody& b=bodies[c]; st.mkb[c]=b.mkbound; st.nid[c]=b.nid;
    for(unsigned i=0;i<3;i++){ st.id[c*3+i]=b.id[i]; st.dis[c*3+i]=b.dis[i]; for(int k=0;k<3;k++)st.ps[(c*3+i)*3+k]=b.ps[i][k]; }
  }
  hd.add_array_ext("MkBound",BT_USHORT,nmk,st.mkb.data()); hd.add_array_ext("Nid",BT_UINT,nmk,st.nid.data());
  hd.add_array_ext("Id",BT_UINT,nmk*3,st.id.data()); hd.add_array_ext("Ps",BT_DOUBLE3,nmk*3,st.ps.data()); hd.add_array_ext("Dis",BT_DOUBLE,nmk*3,st.dis.data());
}
}

void MotionRefSave::save_part(const std::string& dirdata,const std::string& appname,double timeout,unsigned mkboundfirst,bool fresh,
                              unsigned cpart,double timestep,unsigned step,uint32_t n,const uint32_t* idp,const double* pos){
  const unsigned nmk=unsigned(bodies.size()); if(!nmk)return;
  Bi4Item hd("JPartMotRefBi4"); HeadStore st;
  make_head(hd,st,bodies,nmoving,nfloat,appname,true,timeout,mkboundfirst);
  // current positions of the reference particles (JDsPartMotionSave::LoadPosRefCpu, JDsPartMotionSave.cpp:188-208)
  std::vector<double> posref(size_t(nmk)*9,0.0);
  for(unsigned c=0;c<nmk;c++)for(unsigned i=0;i<3;i++){
    const MotRefBody& b=bodies[c];
    if(i>=b.nid)throw std::runtime_error("PartMotionRef: a body has fewer than three distinct reference particles");
    const uint32_t k=find_id(n,idp,b.id[i]);
    if(k>=n)throw std::runtime_error("PartMotionRef: reference particle "+std::to_string(b.id[i])+" is not in the simulation any more");
    for(int d=0;d<3;d++)posref[(c*3+i)*3+d]=pos[size_t(k)*3+d];
  }
  char nm[32]; snprintf(nm,sizeof(nm),"PART_%04u",cpart);
  Bi4Item pt(nm); pt.hide=true;
  pt.set_uint("Cpart",cpart); pt.set_double("TimeStep",timestep); pt.set_uint("Step",step);
  pt.add_array_ext("PosRef",BT_DOUBLE3,nmk*3,posref.data());
  bi4_save_listapp(dirdata+"PartMotionRef.ibi4","JPartMotRefBi4",hd,&pt,fresh);
}

std::vector<uint32_t> MotionRefSave::ref_ids()const{
  std::vector<uint32_t> ids;
  for(const MotRefBody& b:bodies)for(unsigned i=0;i<3;i++){
    if(i>=b.nid)throw std::runtime_error("PartMotionRef: a body has fewer than three distinct reference particles");
    ids.push_back(b.id[i]);
  }
  return ids;
}

void MotionRefSave::save_extra(const std::string& dirdata,const std::string& appname,double timeout2,unsigned mkboundfirst,bool fresh,
                               unsigned cpart,unsigned count,const double* times,const uint32_t* steps,const double* posref){
  const unsigned nmk=unsigned(bodies.size()); if(!nmk || !count)return;
  Bi4Item hd("JPartMotRefBi4"); HeadStore st;
  make_head(hd,st,bodies,nmoving,nfloat,appname,false,timeout2,mkboundfirst);
  char nm[32]; snprintf(nm,sizeof(nm),"PART_%04u",cpart);
  Bi4Item pt(nm); pt.hide=true;
  pt.set_uint("Cpart",cpart);
  if(count==1){ pt.set_double("TimeStep",times[0]); pt.set_uint("Step",steps[0]); }   // MakeBdPartSingle
  else{ pt.set_uint("Count",count); pt.add_array_ext("TimeStep",BT_DOUBLE,count,times); pt.add_array_ext("Step",BT_UINT,count,steps); }   // MakeBdPartArray
  pt.add_array_ext("PosRef",BT_DOUBLE3,count*nmk*3,posref);
  bi4_save_listapp(dirdata+"PartMotionRef2.ibi4","JPartMotRefBi4",hd,&pt,fresh);
}

} // namespace b2s
