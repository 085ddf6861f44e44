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

void MotionRefSave::save_part(const std::string& dirdata,const std::string& appname,double timeout,unsigned mkboundfirst,bool fresh,
                              unsigned cpart,double timestep,unsigned step,uint32_t n,const uint32_t* idp,const double* pos){
  const unsigned nmk=unsigned(bodies.size()); if(!nmk)return;
  // head (JPartMotRefBi4Save::Config, JPartMotRefBi4Save.cpp:149-191)
  Bi4Item hd("JPartMotRefBi4");
  hd.set_text("AppName",appname); hd.set_uint("FormatVer",230729); hd.set_bool("MainFile",true); hd.set_double("TimeOut",timeout);
  { Bi4Value v; v.name="MkBoundFirst"; v.type=BT_USHORT; v.i=0; v.u=mkboundfirst; v.d=0; v.d3[0]=v.d3[1]=v.d3[2]=0; hd.values.push_back(v); }
  hd.set_uint("MkMovingCount",nmoving); hd.set_uint("MkFloatCount",nfloat); hd.set_uint("MkCount",nmk);
  std::vector<uint16_t> mkb(nmk); std::vector<uint32_t> nid(nmk),id(size_t(nmk)*3); std::vector<double> ps(size_t(nmk)*9),dis(size_t(nmk)*3);
  for(unsigned c=0;c<nmk;c++){
    const MotRefBody& b=bodies[c]; mkb[c]=b.mkbound; nid[c]=b.nid;
    for(unsigned i=0;i<3;i++){ id[c*3+i]=b.id[i]; dis[c*3+i]=b.dis[i]; for(int k=0;k<3;k++)ps[(c*3+i)*3+k]=b.ps[i][k]; }
  }
  hd.add_array_ext("MkBound",BT_USHORT,nmk,mkb.data()); hd.add_array_ext("Nid",BT_UINT,nmk,nid.data());
  hd.add_array_ext("Id",BT_UINT,nmk*3,id.data()); hd.add_array_ext("Ps",BT_DOUBLE3,nmk*3,ps.data()); hd.add_array_ext("Dis",BT_DOUBLE,nmk*3,dis.data());
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

} // namespace b2s
