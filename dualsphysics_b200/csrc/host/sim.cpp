// sim.cpp -- C++ host driver: GenCase case in (Case.xml + Case.bi4), Part_XXXX.bi4 out, Run/ComputeStep loop.
//
// Mirrors, for the hot path only, the host side of the reference (all under /root/reference/src/source):
//   JSph::LoadCaseConfig / LoadConfigCtes / LoadConfigParameters / LoadConfigCommands   JSph.cpp:553-1003
//   JCaseEParms (parameters + <simulationdomain>)                                       JCaseEParms.cpp:150-439
//   JCaseParts::ReadXml (particle blocks)                                               JCaseParts.cpp:79-84,274-297,540-567
//   JSphMk (particle code per block)                                                    JSphMk.cpp:95-125
//   JPartsLoad4::LoadParticles + JSph::LoadCaseParticles / ResizeMapLimits              JPartsLoad4.cpp:164-290, JSph.cpp:1445-1476,2151-2198
//   JSphGpuSingle::Run main loop + SaveData                                             JSphGpuSingle.cpp:950-1118
//   JSph::SavePartData (Part_XXXX.bi4, Part_Head.ibi4 omitted values documented in DESIGN.md)  JSph.cpp:3062-3146
// Everything numerical happens behind the C-ABI (include/b2s.h); this file owns no CUDA code.
#include "../../../include/b2s.h"
#include "../../../include/b2s_host.h"
#include "xml_lite.h"
#include "bi4.h"
#include "motionref.h"
#include "../slab_cuts.h"
#include <cmath>
#include <cfloat>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <sstream>
#include <stdexcept>
#include <algorithm>
#include <ctime>
#include <cstdint>
#include <sys/stat.h>

using namespace b2s;

struct PartBlock { int type; unsigned mktype, mk, begin, count; // type: 0 fixed,1 moving,2 floating,3 fluid
  double massbody, masspart, center[3], inertia[9], linvelini[3], angvelini[3]; int transfree[3], rotfree[3]; unsigned refmotion; };
struct MotionCfg { unsigned ref; b2s_motion m; };   // one movement of the flattened chain of a moving block (idbegin/count filled at upload)

// One gauge of <special><gauges> (JGaugeSystem::ReadXml, JDsGaugeSystem.cpp:244-330) with the timing state of JGaugeItem
struct GaugeCfg {
  std::string name; int type;            // 1 velocity, 2 swl, 3 force, 4 maxz
  double height; float distlimit;        // maxz
  unsigned mkbound;                      // force: target block
  double point[3], point0[3], point2[3], pointdp; float masslimit;
  double computedt, computestart, computeend, outputdt, outputstart, outputend; bool output;
  double computenext, outputnext; int id; bool headdone;
  GaugeCfg():type(0),pointdp(0),masslimit(0),computedt(0),computestart(0),computeend(0),outputdt(0),outputstart(0),outputend(0),output(false),computenext(0),outputnext(0),id(-1),headdone(false){}
};

// One zone of <special><damping> (JDsDamping::ReadXml, JDsDamping.cpp:664-681)
struct DampCfg {
  int type;   // 0 dampingzone, 1 dampingbox
  float over, redumax, factor[3];
  double a[3], b[3], c[3], d[3];   // plane: limitmin, limitmax;  box: limitmin ini/end, limitmax ini/end
  bool usedomain; double domain[10]; unsigned directions;
};

struct b2s_sim {
  std::string casepath, casename, dirout, dirdata, err;
  b2s_params prm;
  b2s_ctx* ctx;
  std::vector<PartBlock> blocks;
  std::vector<MotionCfg> motions;
  std::vector<GaugeCfg> gauges;
  std::vector<DampCfg> dampings;
  double posy2d=0;                                                       // Simulate2DPosY
  int variant=-1;                                                       // -b2svariant (-1: the library default, variant 5)
  unsigned piece=0, npiece=1;                                           // slab rank / ranks: PART files are written as pieces (JPartDataBi4.cpp:99-127)
  std::string svextraparts;                                             // -svextraparts / SaveExtraParts: PARTs that get a PartExtra_XXXX.bi4
  uint32_t out_cap=0; bool out_saved=false;                           // capacity of the device buffer of excluded particles; PartOut head written
  unsigned out_total[4]={0,0,0,0};                                    // excluded particles written to PartOut so far, by motive (0: all)
  MotionRefSave motref; bool motref_ready=false;                       // PartMotionRef.ibi4
  // Extra-frequency output of the body data (TimeOutExtra / -toutx; JSph.cpp:789-790,859-870): PartFloatInfo2.ibi4 and PartMotionRef2.ibi4.
  // Records are collected at the end of the steps that reach the schedule and written with the next PART (JSphCpuSingle.cpp:1201-1206).
  double timepartextra=-1, timepartextranext=DBL_MAX;
  struct ExtraStore { std::vector<double> time; std::vector<uint32_t> step; std::vector<double> cen, posref; std::vector<float> f3[10]; void clear(){ time.clear(); step.clear(); cen.clear(); posref.clear(); for(auto& v:f3)v.clear(); } };
  ExtraStore xfloat, xmotion; bool xfloat_fresh=true, xmotion_fresh=true;
  bool initial_saved, head_saved; unsigned partbegin, partbeginfirst; std::string partbegindir; double partbegintime, partbegindtpre;   // restart (-partbegin)
  // run control (JSph members)
  double timemax, timepart, timepartnext; unsigned part, nstepsbreak; bool svallsteps, svposdouble, svdata;
  double casedp, caseh, caseb, caserhop0, casegamma, casemassb, casemassf;
  double caseposmin[3], caseposmax[3];
  unsigned np0;
  std::vector<uint32_t> slab_cuts; uint32_t slab_maxplane;   // slab loading (b2s_sim_load_slab)
  std::string runcode, date; unsigned part_nstep0; double part_time0, part_dtmin, part_dtmax;   // PART header / PartInfo statistics
  b2s_sim():ctx(NULL),initial_saved(false),head_saved(false),partbegin(0),partbeginfirst(0),partbegintime(0),partbegindtpre(0),timemax(0),timepart(0),timepartnext(0),part(0),nstepsbreak(0),svallsteps(false),svposdouble(false),svdata(true),np0(0),slab_maxplane(0),part_nstep0(0),part_time0(0),part_dtmin(DBL_MAX),part_dtmax(0){
    memset(&prm,0,sizeof(prm));
    // RunCode (8 random characters) and Date like JSph::CalcRunCode / fun::GetDateTime
    const time_t now=time(NULL); char buf[64]; strftime(buf,sizeof(buf),"%d-%m-%Y %H:%M:%S",localtime(&now)); date=buf;
    static const char cs[]="abcdefghijklmnopqrstuvwxyz0123456789"; unsigned seed=unsigned(now)^unsigned(uintptr_t(this)>>4);
    for(int k=0;k<8;k++){ seed=seed*1664525u+1013904223u; runcode+=cs[(seed>>16)%36]; }
  }
};

static std::string g_host_err;

// ---- small helpers ------------------------------------------------------------------------------------
static std::string lower(std::string s){ for(auto& c:s)c=char(tolower((unsigned char)c)); return s; }
static std::string upper(std::string s){ for(auto& c:s)c=char(toupper((unsigned char)c)); return s; }
static double todouble(const std::string& s){ return atof(s.c_str()); }
static bool read_xyz(const XmlNode* n,double v[3]){ if(!n)return false; v[0]=todouble(n->attr("x","0")); v[1]=todouble(n->attr("y","0")); v[2]=todouble(n->attr("z","0")); return true; }
static void mkdirs(const std::string& d){ std::string cur; for(size_t i=0;i<d.size();i++){ cur+=d[i]; if(d[i]=='/'||i+1==d.size())mkdir(cur.c_str(),0777); } }

// JCaseEParms::CheckPosValue (JCaseEParms.cpp:374-414): mode 0 default, 1 fixed, 2 default+-value, 3 default+-percent
static int parse_posvalue(const std::string& value,bool isposmin,double& out){
  std::string v; for(char c:value)if(c!=' '&&c!='\t')v+=char(tolower((unsigned char)c));
  out=0;
  if(v=="default"||v.empty())return 0;
  if(v.compare(0,8,"default+")==0 || v.compare(0,8,"default-")==0){
    const char sign=v[7];
    if((isposmin && sign=='+') || (!isposmin && sign=='-'))throw std::runtime_error("simulationdomain: wrong sign in '"+value+"'");
    const bool prc=(!v.empty() && v[v.size()-1]=='%');
    out=atof(v.substr(8,v.size()-(prc? 9: 8)).c_str());
    return prc? 3: 2;
  }
  out=atof(v.c_str());
  return 1;
}

struct EParms {
  std::vector<std::pair<std::string,std::string>> kv;
  std::string posmin[3], posmax[3];
  bool exists(const std::string& k)const{ for(auto& p:kv)if(p.first==k)return true; return false; }
  std::string str(const std::string& k,const std::string& def="")const{ for(auto& p:kv)if(p.first==k)return p.second; return def; }
  double num(const std::string& k,double def)const{ return exists(k)? todouble(str(k)): def; }
  int inum(const std::string& k,int def)const{ return exists(k)? atoi(str(k).c_str()): def; }
};

// ---- case loading --------------------------------------------------------------------------------------
static void load_case(b2s_sim* s,const std::string& options){
  const std::string filexml=s->casepath+".xml";
  std::unique_ptr<XmlNode> root=xml_load(filexml);
  if(root->name!="case")throw std::runtime_error("Case.xml: root element <case> not found");
  const XmlNode* exe=root->child("execution");
  if(!exe)throw std::runtime_error("Case.xml: <execution> not found");
  b2s_params& p=s->prm;
  // -- constants (JCaseCtes::ReadXmlRun)
  const XmlNode* ct=exe->child("constants");
  if(!ct)throw std::runtime_error("Case.xml: <execution><constants> not found");
  auto cval=[&](const char* n,bool opt=false,double def=0){ const XmlNode* e=ct->child(n); if(!e){ if(opt)return def; throw std::runtime_error(std::string("Case.xml: constant missing: ")+n); } return todouble(e->attr("value","0")); };
  { const XmlNode* e=ct->child("data2d"); const std::string v=(e? lower(e->attr("value","false")): "false"); p.simulate2d=(v=="true"||v=="1")? 1: 0; }
  // 2-D (JSph.cpp:646-647): one plane of particles at y=data2dposy; the second component of the acceleration is zeroed
  // (JSphCpuSingle.cpp:717-724) and the kernel constants are the 2-D ones (b2s_config_finalize)
  s->posy2d=(p.simulate2d? cval("data2dposy",true,0): 0);
  double g[3]; if(!read_xyz(ct->child("gravity"),g))throw std::runtime_error("Case.xml: gravity missing");
  for(int k=0;k<3;k++)p.gravity[k]=float(g[k]);
  p.cflnumber=cval("cflnumber"); p.gamma=float(cval("gamma")); p.rhopzero=float(cval("rhop0")); p.dp=cval("dp");
  p.kernelh=float(cval("h")); p.cteb=float(cval("b")); p.massbound=float(cval("massbound")); p.massfluid=float(cval("massfluid"));
  // -- parameters (JCaseEParms + JSph::LoadConfigParameters)
  EParms ep;
  const XmlNode* pr=exe->child("parameters");
  if(!pr)throw std::runtime_error("Case.xml: <execution><parameters> not found");
  for(auto& c:pr->children){
    if(c->name=="parameter")ep.kv.emplace_back(c->attr("key"),c->attr("value"));
    else if(c->name=="simulationdomain"){
      const XmlNode* a=c->child("posmin"); const XmlNode* b=c->child("posmax");
      const char* ax[3]={"x","y","z"};
      for(int k=0;k<3;k++){ ep.posmin[k]=(a? a->attr(ax[k],"default"): "default"); ep.posmax[k]=(b? b->attr(ax[k],"default"): "default"); }
    }
  }
  switch(ep.inum("Kernel",2)){ case 1: p.tkernel=B2S_KERNEL_CUBIC; break; case 2: p.tkernel=B2S_KERNEL_WENDLAND; break; default: throw std::runtime_error("Kernel choice is not valid."); }
  if(ep.inum("RigidAlgorithm",1)!=1)throw std::runtime_error("only RigidAlgorithm=1 (SPH) is supported");
  switch(ep.inum("StepAlgorithm",1)){ case 1: p.tstep=B2S_STEP_VERLET; break; case 2: p.tstep=B2S_STEP_SYMPLECTIC; break; default: throw std::runtime_error("Step algorithm is not valid."); }
  p.verletsteps=ep.inum("VerletSteps",40);
  switch(ep.inum("ViscoTreatment",1)){ case 1: p.tvisco=B2S_VISCO_ARTIFICIAL; break; case 2: p.tvisco=B2S_VISCO_LAMINARSPS; break; case 3: p.tvisco=B2S_VISCO_LAMINAR; break; default: throw std::runtime_error("Viscosity treatment is not valid."); }
  if(!ep.exists("Visco"))throw std::runtime_error("parameter Visco missing");
  p.visco=float(ep.num("Visco",0)); p.viscoboundfactor=float(ep.num("ViscoBoundFactor",1));
  switch(ep.inum("Boundary",1)){ case 1: p.tboundary=B2S_BC_DBC; break; case 2: p.tboundary=B2S_BC_MDBC; break; default: throw std::runtime_error("Boundary Condition method is not valid."); }
  p.slipmode=0;
  if(p.tboundary==B2S_BC_MDBC)p.slipmode=ep.inum("SlipMode",1);
  { const int tddt=ep.inum("DensityDT",0); if(tddt<0||tddt>3)throw std::runtime_error("Density Diffusion Term mode is not valid."); p.tdensity=tddt; p.ddtvalue=float(ep.num("DensityDTvalue",0.1)); }
  { // standard shifting (JSph.cpp:744-758); zones (<special><shifting>) and the advanced mode 4 are rejected below
    const int sm=ep.inum("Shifting",0);
    if(sm<0||sm>4)throw std::runtime_error("Shifting mode in <execution><parameters> is not valid.");
    if(sm==4)throw std::runtime_error("advanced shifting (Shifting=4) is not supported by this build");
    p.shiftmode=sm; p.shiftcoef=float(ep.num("ShiftCoef",-2)); p.shifttfs=float(ep.num("ShiftTFS",0));
  }
  p.ftpause=ep.num("FtPause",0);
  // TimeMax and TimeOut are mandatory in the reference (GetValueDouble without a default, JSph.cpp:786-787)
  if(!ep.exists("TimeMax") || !ep.exists("TimeOut"))throw std::runtime_error("Case.xml: the parameters TimeMax and TimeOut are mandatory");
  s->timemax=ep.num("TimeMax",0); s->timepart=ep.num("TimeOut",0);
  s->timepartextra=ep.num("TimeOutExtra",-1);                            // optional (JSph.cpp:789)
  p.dtini=std::max(0.0,ep.num("DtIni",0)); p.dtmin=std::max(0.0,ep.num("DtMin",0)); p.coefdtmin=double(float(ep.num("CoefDtMin",0.05)));
  p.dt_allparticles=(ep.inum("DtAllParticles",0)==1);
  if(ep.num("DtFixed",0)!=0)throw std::runtime_error("DtFixed is not supported by this build");
  p.rhopoutmin=float(ep.num("RhopOutMin",700)); p.rhopoutmax=float(ep.num("RhopOutMax",1300));
  s->svposdouble=(ep.inum("SavePosDouble",0)!=0);
  s->svextraparts=ep.str("SaveExtraParts","");
  if(s->svextraparts.find_first_not_of("0123456789,-")!=std::string::npos)throw std::runtime_error("SaveExtraParts: a PART interval or a list like 1,4,7-9");
  { // periodic boundaries (JSph.cpp:815-827)
    bool px=false,py=false,pz=false;
    for(int k=0;k<3;k++)p.perixinc[k]=p.periyinc[k]=p.perizinc[k]=0;
    if(ep.exists("XPeriodicIncY")){ p.perixinc[1]=ep.num("XPeriodicIncY",0); px=true; }
    if(ep.exists("XPeriodicIncZ")){ p.perixinc[2]=ep.num("XPeriodicIncZ",0); px=true; }
    if(ep.exists("YPeriodicIncX")){ p.periyinc[0]=ep.num("YPeriodicIncX",0); py=true; }
    if(ep.exists("YPeriodicIncZ")){ p.periyinc[2]=ep.num("YPeriodicIncZ",0); py=true; }
    if(ep.exists("ZPeriodicIncX")){ p.perizinc[0]=ep.num("ZPeriodicIncX",0); pz=true; }
    if(ep.exists("ZPeriodicIncY")){ p.perizinc[1]=ep.num("ZPeriodicIncY",0); pz=true; }
    if(ep.exists("XYPeriodic")){ px=true;  py=true;  pz=false; for(int k=0;k<3;k++)p.perixinc[k]=p.periyinc[k]=p.perizinc[k]=0; }
    if(ep.exists("XZPeriodic")){ px=true;  py=false; pz=true;  for(int k=0;k<3;k++)p.perixinc[k]=p.periyinc[k]=p.perizinc[k]=0; }
    if(ep.exists("YZPeriodic")){ px=false; py=true;  pz=true;  for(int k=0;k<3;k++)p.perixinc[k]=p.periyinc[k]=p.perizinc[k]=0; }
    p.periactive=(px? 1: 0)|(py? 2: 0)|(pz? 4: 0);
  }
  // -- command-line style overrides (JSphCfgRun::LoadArgv + JSph::LoadConfigCommands)
  int cellmode_half=0; int opt_boundary=-1, opt_slip=1;
  {
    std::stringstream ss(options); std::string tok;
    while(ss>>tok){
      if(tok[0]!='-')throw std::runtime_error("invalid option '"+tok+"'");
      std::string w=tok.substr(1), a; size_t c=w.find(':'); if(c!=std::string::npos){ a=w.substr(c+1); w=w.substr(0,c); }
      w=upper(w);
      auto a2=[&](int i){ std::string t=a; for(int k=0;k<i;k++){ size_t q=t.find(':'); t=(q==std::string::npos? "": t.substr(q+1)); } size_t q=t.find(':'); return q==std::string::npos? t: t.substr(0,q); };
      if(w=="GPU"||w=="CPU"||w=="STABLE"||w=="SVRES"||w=="SVTIMERS"||w=="NORTIMES"||w=="OMPTHREADS"||w=="SVDOMAINVTK"){}
      else if(w=="SYMPLECTIC")p.tstep=B2S_STEP_SYMPLECTIC;
      else if(w=="VERLET"){ p.tstep=B2S_STEP_VERLET; if(!a.empty())p.verletsteps=atoi(a.c_str()); }
      else if(w=="WENDLAND")p.tkernel=B2S_KERNEL_WENDLAND;
      else if(w=="CUBIC")p.tkernel=B2S_KERNEL_CUBIC;
      else if(w=="VISCOART"){ p.tvisco=B2S_VISCO_ARTIFICIAL; p.visco=float(atof(a.c_str())); }
      else if(w=="VISCOLAM"){ p.tvisco=B2S_VISCO_LAMINAR; p.visco=float(atof(a.c_str())); }
      else if(w=="VISCOLAMSPS"){ p.tvisco=B2S_VISCO_LAMINARSPS; p.visco=float(atof(a.c_str())); }
      else if(w=="VISCOBOUNDFACTOR")p.viscoboundfactor=float(atof(a.c_str()));
      else if(w=="DDT"){ p.tdensity=atoi(a.c_str()); if(p.tdensity<0||p.tdensity>3)throw std::runtime_error("Density Diffusion Term mode is not valid."); if(p.ddtvalue==0)p.ddtvalue=0.1f; }
      else if(w=="DDTVALUE")p.ddtvalue=float(atof(a.c_str()));
      else if(w=="DBC"){ opt_boundary=1; }
      else if(w=="MDBC"){ opt_boundary=2; opt_slip=1; }
      else if(w=="MDBC_NOSLIP"){ opt_boundary=2; opt_slip=2; }
      else if(w=="MDBC_FREESLIP"){ opt_boundary=2; opt_slip=3; }
      else if(w=="CELLMODE"){ const std::string u=upper(a); if(u=="HALF"||u=="H")cellmode_half=1; else if(u=="FULL"||u=="2H")cellmode_half=0; else throw std::runtime_error("invalid cellmode"); }
      else if(w=="RHOPOUT"){ p.rhopoutmin=float(atof(a2(0).c_str())); p.rhopoutmax=float(atof(a2(1).c_str())); }
      else if(w=="SHIFTING"){   // -shifting:<mode> with default parameters (JSphCfgRun.cpp:423-436, JSph.cpp:939-953)
        const std::string u=upper(a);
        if(u=="NONE")p.shiftmode=B2S_SHIFT_NONE; else if(u=="NOBOUND")p.shiftmode=B2S_SHIFT_NOBOUND; else if(u=="NOFIXED")p.shiftmode=B2S_SHIFT_NOFIXED; else if(u=="FULL")p.shiftmode=B2S_SHIFT_FULL;
        else throw std::runtime_error("shifting mode '"+a+"' is not supported by this build (none, nobound, nofixed, full)");
        p.shiftcoef=-2.f; p.shifttfs=0.f;
      }
      else if(w=="CFL")p.cflnumber=atof(a.c_str());
      else if(w=="FTPAUSE")p.ftpause=atof(a.c_str());
      else if(w=="TMAX")s->timemax=atof(a.c_str());
      else if(w=="TOUT")s->timepart=atof(a.c_str());
      else if(w=="TOUTX")s->timepartextra=atof(a.c_str());                // JSphCfgRun.cpp:494-496
      else if(w=="NSTEPS")s->nstepsbreak=unsigned(atoi(a.c_str()));
      else if(w=="PARTBEGIN"){   // -partbegin:begin[:first] dir  (JSphCfgRun.cpp:462-470): restart from dir/Part_<begin>.bi4
        s->partbegin=unsigned(atoi(a2(0).c_str())); s->partbeginfirst=(a2(1).empty()? s->partbegin: unsigned(atoi(a2(1).c_str())));
        if(!(ss>>s->partbegindir) || s->partbegindir[0]=='-')throw std::runtime_error("-partbegin needs the directory of the PART files as next argument");
        if(!s->partbegin)throw std::runtime_error("-partbegin:0 is not a restart");
      }
      else if(w=="SVSTEPS")s->svallsteps=(a.empty()||atoi(a.c_str())!=0);
      else if(w=="SVEXTRAPARTS"){ if(a.empty() || a.find_first_not_of("0123456789,-")!=std::string::npos)throw std::runtime_error("-svextraparts: a PART interval or a list like 1,4,7-9"); s->svextraparts=a; }
      else if(w=="SAVEPOSDOUBLE")s->svposdouble=(a.empty()||atoi(a.c_str())!=0);
      else if(w=="B2SVARIANT"){   // interaction kernel (include/b2s.h b2s_set_interaction_variant): 5 = packed FP32, fastest (default); 0 = reference summation order
        s->variant=atoi(a.c_str()); if(a.empty() || (s->variant!=0 && s->variant!=5 && s->variant!=6))throw std::runtime_error("-b2svariant: 0, 5 or 6");
      }
      else if(w=="SV"){ s->svdata=(lower(a).find("none")==std::string::npos); }
      else throw std::runtime_error("option not supported by this build: "+tok);
    }
  }
  if(opt_boundary>0){ p.tboundary=(opt_boundary==2? B2S_BC_MDBC: B2S_BC_DBC); p.slipmode=(opt_boundary==2? opt_slip: 0); }
  // after the -tmax / -tout overrides: a zero TimeOut would never advance the PART schedule (JDsOutputTime needs a positive step)
  if(!(s->timemax>0))throw std::runtime_error("TimeMax (or -tmax) must be greater than zero");
  if(!(s->timepart>0) && !s->svallsteps)throw std::runtime_error("TimeOut (or -tout) must be greater than zero");
  if(p.tboundary==B2S_BC_MDBC && p.slipmode>=B2S_SLIP_NOSLIP && p.viscoboundfactor!=1.f)throw std::runtime_error("ViscoBoundFactor must be 1 with mDBC2 (JSph.cpp:1603)");
  if(p.tboundary==B2S_BC_MDBC && (p.slipmode<1 || p.slipmode>3))throw std::runtime_error("mDBC slip mode is invalid; use SlipMode=1");
  p.mdbc_corrector=(p.slipmode>=B2S_SLIP_NOSLIP)? 1: 0;   // MdbcCorrector (JSph.cpp:899)
  if(p.tdensity==B2S_DDT_NONE)p.ddtvalue=0;
  p.cellmode_half=cellmode_half;
  if(!(p.rhopoutmin<p.rhopoutmax)){ p.rhopoutmin=-FLT_MAX; p.rhopoutmax=FLT_MAX; }
  // -- particles (JCaseParts::ReadXml)
  const XmlNode* pa=exe->child("particles");
  if(!pa)throw std::runtime_error("Case.xml: <execution><particles> not found");
  unsigned cnt[4]={0,0,0,0};
  for(auto& c:pa->children){
    int t=-1;
    if(c->name=="fixed")t=0; else if(c->name=="moving")t=1; else if(c->name=="floating")t=2; else if(c->name=="fluid")t=3;
    else if(c->name=="properties"||c->name[0]=='_')continue;
    else throw std::runtime_error("Case.xml: unknown particle block <"+c->name+">");
    PartBlock b; memset(&b,0,sizeof(b));
    b.type=t; b.mktype=unsigned(atoi(c->attr(t==3? "mkfluid": "mkbound","0").c_str())); b.mk=unsigned(atoi(c->attr("mk","0").c_str()));
    b.begin=unsigned(strtoul(c->attr("begin","0").c_str(),NULL,10)); b.count=unsigned(strtoul(c->attr("count","0").c_str(),NULL,10));
    if(t==1)b.refmotion=unsigned(atoi(c->attr("refmotion","0").c_str()));
    if(t==2){
      const XmlNode* e;
      if(!(e=c->child("massbody")))throw std::runtime_error("floating: massbody missing");
      b.massbody=todouble(e->attr("value"));
      if(!(e=c->child("masspart")))throw std::runtime_error("floating: masspart missing");
      b.masspart=todouble(e->attr("value"));
      if(!read_xyz(c->child("center"),b.center))throw std::runtime_error("floating: center missing");
      e=c->child("inertia"); if(!e)throw std::runtime_error("floating: inertia missing");
      if(e->has("x")&&e->has("y")&&e->has("z")){ double v[3]; read_xyz(e,v); b.inertia[0]=v[0]; b.inertia[4]=v[1]; b.inertia[8]=v[2]; }
      else{
        for(int r=1;r<=3;r++)for(int q=1;q<=3;q++){
          char nm[8]; snprintf(nm,sizeof(nm),"v%d%d",r,q);
          if(e->has(nm))b.inertia[(r-1)*3+q-1]=todouble(e->attr(nm));
          for(const XmlNode* row:e->childs("values"))if(row->has(nm))b.inertia[(r-1)*3+q-1]=todouble(row->attr(nm));
        }
      }
      for(int k=0;k<3;k++){ b.transfree[k]=1; b.rotfree[k]=1; }
      const XmlNode* tf=(c->child("translation")? c->child("translation"): c->child("translationDOF"));
      const XmlNode* rf=(c->child("rotation")? c->child("rotation"): c->child("rotationDOF"));
      double v[3];
      if(tf && read_xyz(tf,v))for(int k=0;k<3;k++)b.transfree[k]=(v[k]!=0);
      if(rf && read_xyz(rf,v))for(int k=0;k<3;k++)b.rotfree[k]=(v[k]!=0);
      const XmlNode* lv=(c->child("linearvelini")? c->child("linearvelini"): c->child("velini"));
      const XmlNode* av=(c->child("angularvelini")? c->child("angularvelini"): c->child("omegaini"));
      if(lv)read_xyz(lv,b.linvelini);
      if(av)read_xyz(av,b.angvelini);
      if(c->child("linearvel")||c->child("angularvel")||c->child("linearforce")||c->child("angularforce"))
        throw std::runtime_error("floating: imposed velocity/force is not supported by this build");
    }
    cnt[t]+=b.count;
    s->blocks.push_back(b);
  }
  p.casenfixed=cnt[0]; p.casenmoving=cnt[1]; p.casenfloat=cnt[2]; p.casenfluid=cnt[3]; p.casenp=cnt[0]+cnt[1]+cnt[2]+cnt[3];
  if(p.tboundary==B2S_BC_MDBC && p.slipmode>=B2S_SLIP_NOSLIP && p.casenfloat)throw std::runtime_error("mDBC2 (SlipMode 2/3) with floating bodies is not supported by this build");
  // 2-D: the ghost-node systems of mDBC are 3x3 there (JSphCpu_mdbc.cpp:158,448) and a floating body is locked to the plane
  // (JSph.cpp:2629-2640): neither is built
  if(p.simulate2d && p.tboundary==B2S_BC_MDBC)throw std::runtime_error("2-D cases with mDBC are not supported by this build");
  if(p.simulate2d && p.casenfloat)throw std::runtime_error("2-D cases with floating bodies are not supported by this build");
  if(p.simulate2d && (p.periactive&2))throw std::runtime_error("Cannot use periodic conditions in Y with 2D simulations");
  { const unsigned np=unsigned(strtoul(pa->attr("np","0").c_str(),NULL,10)), nb=unsigned(strtoul(pa->attr("nb","0").c_str(),NULL,10)), nbf=unsigned(strtoul(pa->attr("nbf","0").c_str(),NULL,10));
    if(np!=p.casenp || nb!=p.casenp-p.casenfluid || nbf!=p.casenfixed)throw std::runtime_error("The amount of particles does not match the header."); }
  // -- gauges (JGaugeSystem::LoadXml "case.execution.special.gauges"): velocity and swl
  if(const XmlNode* sp=exe->child("special")){
    if(sp->child("shifting"))throw std::runtime_error("<special><shifting> (shifting zones) is not supported by this build; use the Shifting parameters");
    if(const XmlNode* ds=sp->child("damping")){
      const std::string dact=lower(ds->attr("active","true"));
      if(!(dact=="false"||dact=="0"))for(auto& cu:ds->children){
        const XmlNode* c=cu.get();
        if(c->name.empty() || c->name[0]=='_')continue;
        const std::string act=lower(c->attr("active","true"));
        if(act=="false"||act=="0")continue;
        DampCfg d; memset(&d,0,sizeof(d));
        auto fval=[&](const char* n,const char* a,bool opt,double def){ const XmlNode* e=c->child(n); if(!e||!e->has(a)){ if(opt)return def; throw std::runtime_error("damping: <"+std::string(n)+" "+a+"> missing"); } return todouble(e->attr(a)); };
        d.over=float(fval("overlimit","value",false,0)); d.redumax=float(fval("redumax","value",true,10));
        d.factor[0]=float(fval("factorxyz","x",true,1)); d.factor[1]=float(fval("factorxyz","y",true,1)); d.factor[2]=float(fval("factorxyz","z",true,1));
        if(c->name=="dampingzone"){
          d.type=0;
          if(!read_xyz(c->child("limitmin"),d.a)||!read_xyz(c->child("limitmax"),d.b))throw std::runtime_error("dampingzone: limitmin/limitmax missing");
          if(const XmlNode* dom=c->child("domain")){
            d.usedomain=true;
            if(!dom->has("zmin")||!dom->has("zmax"))throw std::runtime_error("dampingzone: <domain zmin zmax> missing");
            d.domain[0]=todouble(dom->attr("zmin")); d.domain[1]=todouble(dom->attr("zmax"));
            const char* pn[4]={"point1","point2","point3","point4"};
            for(int k=0;k<4;k++){ const XmlNode* q=dom->child(pn[k]); if(!q||!q->has("x")||!q->has("y"))throw std::runtime_error("dampingzone: <domain> point missing"); d.domain[2+2*k]=todouble(q->attr("x")); d.domain[3+2*k]=todouble(q->attr("y")); }
          }
        }
        else if(c->name=="dampingbox"){
          d.type=1;
          const XmlNode* emin=c->child("limitmin"); const XmlNode* emax=c->child("limitmax");
          if(!emin||!emax||!read_xyz(emin->child("pointini"),d.a)||!read_xyz(emin->child("pointend"),d.b)||!read_xyz(emax->child("pointini"),d.c)||!read_xyz(emax->child("pointend"),d.d))
            throw std::runtime_error("dampingbox: limitmin/limitmax pointini/pointend missing");
          // directions (JDsDampingOp_Box::ReadXmlDirections, JDsDamping.cpp:313-340): list of [+-]name, default all
          d.directions=B2S_DAMPDIR_ALL;
          std::string dirs=(c->child("directions")? lower(c->child("directions")->attr("value","")): "");
          for(auto& ch:dirs)if(ch==',')ch=' ';
          std::stringstream ss(dirs); std::string v; bool any=false;
          while(ss>>v){
            if(!any){ d.directions=0; any=true; }
            bool add=true;
            if(v[0]=='+')v=v.substr(1);
            if(!v.empty() && v[0]=='-'){ v=v.substr(1); add=false; }
            unsigned bd=0;
            if(v=="top")bd=B2S_DAMPDIR_TOP; else if(v=="bottom")bd=B2S_DAMPDIR_BOTTOM; else if(v=="left")bd=B2S_DAMPDIR_LEFT; else if(v=="right")bd=B2S_DAMPDIR_RIGHT;
            else if(v=="front")bd=B2S_DAMPDIR_FRONT; else if(v=="back")bd=B2S_DAMPDIR_BACK; else if(v=="all")bd=B2S_DAMPDIR_ALL;
            else throw std::runtime_error("dampingbox: <directions>: some value is invalid");
            if(add)d.directions|=bd; else d.directions&=(B2S_DAMPDIR_ALL^bd);
          }
          if(!d.directions)throw std::runtime_error("dampingbox: at least one direction must be configured");
        }
        else throw std::runtime_error("damping <"+c->name+"> is not supported by this build (dampingzone, dampingbox)");
        s->dampings.push_back(d);
      }
    }
    if(const XmlNode* gs=sp->child("gauges")){
      struct Def { double cdt,cs,ce,odt,os,oe; bool out; } def={-1,0,-1,-1,0,-1,false};   // <0: TimeOut / TimeMax, resolved after the command-line options
      auto common=[&](const XmlNode* e,Def d)->Def{
        if(!e)return d;
        if(const XmlNode* c=e->child("computedt"))d.cdt=todouble(c->attr("value","0"));
        if(const XmlNode* c=e->child("computetime")){ if(c->has("start"))d.cs=todouble(c->attr("start")); if(c->has("end"))d.ce=todouble(c->attr("end")); }
        if(const XmlNode* c=e->child("output")){ const std::string v=lower(c->attr("value","false")); d.out=(v=="true"||v=="1"); }
        if(const XmlNode* c=e->child("outputdt"))d.odt=todouble(c->attr("value","0"));
        if(const XmlNode* c=e->child("outputtime")){ if(c->has("start"))d.os=todouble(c->attr("start")); if(c->has("end"))d.oe=todouble(c->attr("end")); }
        return d;
      };
      def=common(gs->child("default"),def);
      for(auto& cu:gs->children){
        const XmlNode* c=cu.get();
        if(c->name.empty() || c->name[0]=='_' || c->name=="default")continue;
        const std::string act=lower(c->attr("active","true"));
        if(act=="false"||act=="0")continue;
        GaugeCfg g; g.name=c->attr("name");
        if(g.name.empty())throw std::runtime_error("gauge without name");
        for(auto& o:s->gauges)if(o.name==g.name)throw std::runtime_error("The gauge name '"+g.name+"' already exists.");
        const Def d=common(c,def);
        g.computedt=d.cdt; g.computestart=d.cs; g.computeend=d.ce; g.outputdt=d.odt; g.outputstart=d.os; g.outputend=d.oe; g.output=d.out;
        if(c->name=="velocity"){
          g.type=1;
          if(!read_xyz(c->child("point"),g.point))throw std::runtime_error("gauge velocity: <point> missing");
        }
        else if(c->name=="swl"){
          g.type=2;
          if(!read_xyz(c->child("point0"),g.point0)||!read_xyz(c->child("point2"),g.point2))throw std::runtime_error("gauge swl: point0/point2 missing");
          const XmlNode* pd=c->child("pointdp");
          if(pd && pd->has("coefdp"))g.pointdp=p.dp*atof(pd->attr("coefdp").c_str());
          else if(pd)g.pointdp=double(float(atof(pd->attr("value","0").c_str())));
          if(g.pointdp<=0)throw std::runtime_error("gauge swl: the pointdp is invalid");
          const XmlNode* ml=c->child("masslimit");
          if(ml && ml->has("value"))g.masslimit=float(atof(ml->attr("value").c_str()));
          else if(ml && ml->has("coef"))g.masslimit=p.massfluid*float(atof(ml->attr("coef").c_str()));
          else g.masslimit=p.massfluid*(p.simulate2d? 0.4f: 0.5f);
          if(g.masslimit<=0)throw std::runtime_error("gauge swl: the masslimit is invalid");
        }
        else if(c->name=="force"){
          g.type=3;
          const XmlNode* tg=c->child("target");
          if(!tg || !tg->has("mkbound"))throw std::runtime_error("gauge force: <target mkbound> missing");
          g.mkbound=unsigned(atoi(tg->attr("mkbound").c_str()));
        }
        else if(c->name=="maxz"){
          g.type=4;
          if(!read_xyz(c->child("point0"),g.point0))throw std::runtime_error("gauge maxz: <point0> missing");
          const XmlNode* hh=c->child("height");
          if(!hh || !hh->has("value"))throw std::runtime_error("gauge maxz: <height value> missing");
          g.height=double(float(atof(hh->attr("value").c_str())));
          const XmlNode* dl=c->child("distlimit");
          g.distlimit=0;
          if(dl && dl->has("value"))g.distlimit=float(atof(dl->attr("value").c_str()));
          else if(dl && dl->has("coefdp"))g.distlimit=float(p.dp*double(float(atof(dl->attr("coefdp").c_str()))));
          else if(dl && dl->has("coefh"))g.distlimit=float(double(p.kernelh)*double(float(atof(dl->attr("coefh").c_str()))));
          if(g.distlimit<=0)throw std::runtime_error("gauge maxz: the distlimit is invalid");
        }
        else throw std::runtime_error("gauge <"+c->name+"> is not supported by this build (velocity, swl, maxz, force)");
        s->gauges.push_back(g);
      }
    }
  }
  // -- motion (JMotion::ReadXml, JMotion.cpp:692-903): <objreal ref> with wait / mvrect / mvrectace / mvrot / mvrotace / mvrectsinu /
  //    mvrotsinu, chained through next=; every <begin mov start [finish]> starts a chain that is flattened here into time windows
  if(const XmlNode* mo=exe->child("motion")){
    for(const XmlNode* o:mo->childs("objreal")){
      const unsigned ref=unsigned(atoi(o->attr("ref","0").c_str()));
      if(!o->childs("obj").empty() || !o->childs("objreal").empty())throw std::runtime_error("motion: nested objects are not supported by this build");
      struct MovDef { int id,next; double duration; b2s_motion m; };
      std::vector<MovDef> defs;
      struct Begin { int mov; double start,finish; };
      std::vector<Begin> begins;
      for(auto& c:o->children){
        if(c->name.empty() || c->name[0]=='_')continue;
        if(c->name=="begin"){ begins.push_back({atoi(c->attr("mov","0").c_str()),todouble(c->attr("start","0")),(c->has("finish")? todouble(c->attr("finish")): -1.0)}); continue; }
        MovDef d; memset(&d.m,0,sizeof(d.m));
        d.id=atoi(c->attr("id","0").c_str()); d.next=atoi(c->attr("next","0").c_str()); d.duration=todouble(c->attr("duration","0"));
        const std::string units=c->attr("anglesunits","degrees");
        if(units!="degrees" && units!="radians")throw std::runtime_error("motion: anglesunits must be degrees or radians");
        const bool deg=(units=="degrees");
        const double todeg=(deg? 1.0: 180.0/M_PI), torad=(deg? M_PI/180.0: 1.0);
        auto ang=[&](const char* n,const char* at,bool req,double def){ const XmlNode* e=c->child(n); if(!e||!e->has(at)){ if(req)throw std::runtime_error("motion <"+c->name+">: <"+n+" "+at+"> missing"); return def; } return todouble(e->attr(at)); };
        auto axis=[&](){ if(!read_xyz(c->child("axisp1"),d.m.axisp1)||!read_xyz(c->child("axisp2"),d.m.axisp2))throw std::runtime_error("motion <"+c->name+">: axisp1/axisp2 missing"); };
        if(c->name=="wait")d.m.type=B2S_MOV_WAIT;
        else if(c->name=="mvrect"){ d.m.type=B2S_MOV_RECT; if(!read_xyz(c->child("vel"),d.m.v1))throw std::runtime_error("mvrect: <vel> missing"); }
        else if(c->name=="mvrectace"){
          d.m.type=B2S_MOV_RECTACE;
          if(!read_xyz(c->child("ace"),d.m.v2))throw std::runtime_error("mvrectace: <ace> missing");
          if(!read_xyz(c->child("velini"),d.m.v1))throw std::runtime_error("mvrectace without <velini> (velocity of the previous movement) is not supported by this build");
        }
        else if(c->name=="mvrot"){ d.m.type=B2S_MOV_ROT; axis(); d.m.v1[0]=ang("vel","ang",true,0)*todeg; }
        else if(c->name=="mvrotace"){
          d.m.type=B2S_MOV_ROTACE; axis(); d.m.v2[0]=ang("ace","ang",true,0)*todeg;
          if(!c->child("velini"))throw std::runtime_error("mvrotace without <velini> (velocity of the previous movement) is not supported by this build");
          d.m.v1[0]=ang("velini","ang",true,0)*todeg;
        }
        else if(c->name=="mvrectsinu"){
          d.m.type=B2S_MOV_RECTSINU;
          if(!read_xyz(c->child("freq"),d.m.v1)||!read_xyz(c->child("ampl"),d.m.v2))throw std::runtime_error("mvrectsinu: freq/ampl missing");
          if(c->child("phase"))read_xyz(c->child("phase"),d.m.v3); else d.m.pad=1;   // no <phase>: PhasePrev (checked below)
          for(int k=0;k<3;k++)d.m.v3[k]*=torad;
        }
        else if(c->name=="mvrotsinu"){
          d.m.type=B2S_MOV_ROTSINU; axis();
          d.m.v1[0]=ang("freq","v",true,0); d.m.v2[0]=ang("ampl","v",true,0)*todeg; d.m.v3[0]=ang("phase","v",false,0)*torad;
          if(!c->child("phase"))d.m.pad=1;
        }
        else throw std::runtime_error("motion <"+c->name+"> is not supported by this build (wait, mvrect, mvrectace, mvrot, mvrotace, mvrectsinu, mvrotsinu)");
        defs.push_back(d);
      }
      // a sinusoidal movement without <phase> continues with the phase of the movement before it (JMotionMovActive::NextMov,
      // JMotionObj.cpp:262-280): only the first movement of a chain may omit it here (it then starts at phase 0 like the reference)
      for(auto& q:defs)if(q.m.pad){
        for(auto& o2:defs)if(o2.next==q.id)throw std::runtime_error("motion: a sinusoidal movement without <phase> after another movement (phase of the previous movement) is not supported by this build");
        q.m.pad=0;
      }
      const double tlimit=(s->timemax>0? s->timemax: 1e30);
      for(auto& b:begins){
        double t=b.start; int id=b.mov; size_t guard=0;
        while(id && t<=tlimit && guard++<20000){
          const MovDef* d=NULL; for(auto& q:defs)if(q.id==id)d=&q;
          if(!d)throw std::runtime_error("motion: movement id not found in the object");
          if(!(d->duration>0))break;
          MotionCfg m; m.ref=ref; m.m=d->m; m.m.tstart=t; m.m.tend=t+d->duration;
          if(b.finish>=0 && m.m.tend>b.finish)m.m.tend=b.finish;
          if(m.m.tend>m.m.tstart)s->motions.push_back(m);
          if(b.finish>=0 && t+d->duration>=b.finish)break;
          t+=d->duration; id=d->next;
        }
      }
    }
  }
}

// Device buffer of the excluded particles (drained at every PART): the whole fluid on one GPU, twice its share on a slab rank.
static void enable_out(b2s_sim* s,int nranks){
  if(s->out_cap)return;
  const b2s_params& p=s->prm;
  s->out_cap=(nranks>1? std::max(std::min(p.casenfluid,2u*(p.casenfluid/unsigned(nranks))+65536u),1024u): std::max(p.casenfluid,1024u));
  if(b2s_out_enable(s->ctx,s->out_cap))throw std::runtime_error(std::string("b2s_out_enable failed: ")+b2s_last_error(s->ctx));
}

static void load_particles_and_create(b2s_sim* s,int device,unsigned extra_capacity,int slab_rank=0,int slab_nranks=0){
  b2s_params& p=s->prm;
  std::unique_ptr<Bi4Item> data=bi4_load(s->casepath+".bi4");
  if(data->name!="JPartDataBi4")throw std::runtime_error("Case.bi4: root item JPartDataBi4 not found");
  const Bi4Item* part=data->item("PART_0000");
  if(!part)throw std::runtime_error("Case.bi4: PART_0000 not found");
  if(data->get_uint("CaseNp")!=p.casenp || data->get_uint("CaseNfixed")!=p.casenfixed || data->get_uint("CaseNmoving")!=p.casenmoving
     || data->get_uint("CaseNfloat")!=p.casenfloat || data->get_uint("CaseNfluid")!=p.casenfluid)
    throw std::runtime_error("Particle number does not match the configuration of the case.");
  if(data->get_bool("Data2d")!=(p.simulate2d!=0))throw std::runtime_error("Data file does not match the dimension of the case (2D/3D).");
  unsigned np=unsigned(part->get_uint("Npok"));
  const Bi4Array* aid=part->array("Idp"); const Bi4Array* apd=part->array("Posd"); const Bi4Array* apf=part->array("Pos");
  const Bi4Array* avel=part->array("Vel"); const Bi4Array* arho=part->array("Rhop"); const Bi4Array* anor=part->array("BoundNor");
  if(!aid)throw std::runtime_error("Only Idp (32 bits) is valid at the moment.");
  if((!apd&&!apf)||!avel||!arho)throw std::runtime_error("Case.bi4: particle arrays missing");
  if(aid->count!=np||avel->count!=np||arho->count!=np)throw std::runtime_error("Case.bi4: array sizes do not match Npok");
  const uint32_t* idp=(const uint32_t*)aid->ptr();
  std::vector<double> pos(size_t(np)*3);
  if(apd)memcpy(pos.data(),apd->ptr(),size_t(np)*24);
  else{ const float* f=(const float*)apf->ptr(); for(size_t i=0;i<size_t(np)*3;i++)pos[i]=double(f[i]); }
  const float* vel=(const float*)avel->ptr(); const float* rho=(const float*)arho->ptr();
  std::vector<float> velrho(size_t(np)*4);
  for(unsigned i=0;i<np;i++){ velrho[size_t(i)*4]=vel[size_t(i)*3]; velrho[size_t(i)*4+1]=vel[size_t(i)*3+1]; velrho[size_t(i)*4+2]=vel[size_t(i)*3+2]; velrho[size_t(i)*4+3]=rho[i]; }
  // radius of every floating body = max distance particle-centre (JSph::InitFloatings), from the FULL particle set of Case.bi4,
  // i.e. before a restart replaces the positions: the body's initial centre belongs to the initial positions
  std::vector<float> ftradius;
  for(auto& b:s->blocks)if(b.type==2){
    double r2max=0;
    for(unsigned i=0;i<np;i++)if(idp[i]>=b.begin && idp[i]<b.begin+b.count){ double d2=0; for(int k=0;k<3;k++){ const double d=pos[size_t(i)*3+k]-b.center[k]; d2+=d*d; } r2max=std::max(r2max,d2); }
    ftradius.push_back(float(sqrt(r2max)));
  }
  // Restart (JPartsLoad4::LoadParticles with partbegin, JPartsLoad4.cpp:181-224): particle data of dir/Part_<begin>.bi4 replace the
  // initial state; the case definition (counts, constants, mk blocks) still comes from Case.xml/Case.bi4.
  std::unique_ptr<Bi4Item> pdata;
  std::vector<uint32_t> pb_idp; unsigned np_case=np;
  if(s->partbegin){
    std::string dir=s->partbegindir; if(!dir.empty() && dir[dir.size()-1]!='/')dir+='/';
    char fn[64]; snprintf(fn,sizeof(fn),"Part_%04u.bi4",s->partbegin);
    pdata=bi4_load(dir+fn);
    char pn[32]; snprintf(pn,sizeof(pn),"PART_%04u",s->partbegin);
    const Bi4Item* pp=pdata->item(pn);
    if(!pp)throw std::runtime_error(std::string("restart: item ")+pn+" not found in "+dir+fn);
    if(pdata->get_uint("CaseNp")!=p.casenp || pdata->get_uint("CaseNfluid")!=p.casenfluid)throw std::runtime_error("restart: the PART file does not belong to this case (particle counts differ)");
    const unsigned npp=unsigned(pp->get_uint("Npok"));
    const Bi4Array* qid=pp->array("Idp"); const Bi4Array* qpd=pp->array("Posd"); const Bi4Array* qpf=pp->array("Pos");
    const Bi4Array* qv=pp->array("Vel"); const Bi4Array* qr=pp->array("Rhop");
    if(!qid||(!qpd&&!qpf)||!qv||!qr||qid->count!=npp)throw std::runtime_error("restart: particle arrays missing in the PART file");
    // particles by id (boundary first, as the solver needs them)
    std::vector<unsigned> order(npp); for(unsigned i=0;i<npp;i++)order[i]=i;
    const uint32_t* rid=(const uint32_t*)qid->ptr();
    std::sort(order.begin(),order.end(),[&](unsigned a,unsigned b){ return rid[a]<rid[b]; });
    pb_idp.resize(npp); pos.resize(size_t(npp)*3); velrho.resize(size_t(npp)*4);
    const float* rv=(const float*)qv->ptr(); const float* rr=(const float*)qr->ptr();
    for(unsigned k=0;k<npp;k++){
      const unsigned i=order[k];
      pb_idp[k]=rid[i];
      for(int c=0;c<3;c++)pos[size_t(k)*3+c]=(qpd? ((const double*)qpd->ptr())[size_t(i)*3+c]: double(((const float*)qpf->ptr())[size_t(i)*3+c]));
      for(int c=0;c<3;c++)velrho[size_t(k)*4+c]=rv[size_t(i)*3+c];
      velrho[size_t(k)*4+3]=rr[i];
    }
    s->partbegintime=pp->get_double("TimeStep",0); s->partbegindtpre=pp->get_double("SymplecticDtPre",0);
    np=npp; idp=pb_idp.data();
  }
  (void)np_case;
  // CheckSortParticles: boundary ids occupy the first slots (JPartsLoad4.cpp:131-140)
  const unsigned nbound=p.casenfixed+p.casenmoving+p.casenfloat;
  { unsigned lastbound=0; for(unsigned i=0;i<np;i++)if(idp[i]<nbound && i>lastbound)lastbound=i; if(nbound && lastbound+1!=nbound)throw std::runtime_error("Order of boundary (fixed, moving and floating) particles is invalid."); }
  // CheckRhoLimits (JSph.cpp:2136-2145)
  for(unsigned i=0;i<np;i++)if(idp[i]>=nbound && (velrho[size_t(i)*4+3]<p.rhopoutmin || velrho[size_t(i)*4+3]>p.rhopoutmax))throw std::runtime_error("Initial fluid density is out of limits.");
  // particle codes (JSphMk)
  std::vector<uint16_t> code(np);
  {
    unsigned idx[4]={0,0,0,0};
    const uint16_t tcode[4]={0x0000,0x0800,0x1000,0x1800};
    std::vector<uint16_t> codeofid(p.casenp,0);
    for(auto& b:s->blocks){ const uint16_t c=uint16_t(tcode[b.type]|idx[b.type]); idx[b.type]++; for(unsigned k=0;k<b.count;k++){ if(b.begin+k>=p.casenp)throw std::runtime_error("particle block exceeds np"); codeofid[b.begin+k]=c; } }
    for(unsigned i=0;i<np;i++){ if(idp[i]>=p.casenp)throw std::runtime_error("particle id exceeds np"); code[i]=codeofid[idp[i]]; }
  }
  // map limits: CalculeLimits(KernelH*BORDER_MAP) + ResizeMapLimits (JSph.cpp:2174-2178,1445-1476)
  double cmin[3],cmax[3];
  if(!data->get_double3("CasePosMin",cmin) || !data->get_double3("CasePosMax",cmax) || (cmin[0]==cmax[0]&&cmin[1]==cmax[1]&&cmin[2]==cmax[2])){
    for(int k=0;k<3;k++){ cmin[k]=DBL_MAX; cmax[k]=-DBL_MAX; }
    for(unsigned i=0;i<np;i++)for(int k=0;k<3;k++){ cmin[k]=std::min(cmin[k],pos[size_t(i)*3+k]); cmax[k]=std::max(cmax[k],pos[size_t(i)*3+k]); }
  }
  for(int k=0;k<3;k++){ s->caseposmin[k]=cmin[k]; s->caseposmax[k]=cmax[k]; }
  double mpmin[3],mpmax[3]; bool mapsize=false;
  if(data->get_double3("MapPosMin",mpmin) && data->get_double3("MapPosMax",mpmax))mapsize=!(mpmin[0]==mpmax[0]&&mpmin[1]==mpmax[1]&&mpmin[2]==mpmax[2]);
  if(mapsize){ for(int k=0;k<3;k++){ p.maprealposmin[k]=mpmin[k]; p.maprealposmax[k]=mpmax[k]; } }
  else{
    const double border=double(p.kernelh)*0.05;   // BORDER_MAP (DualSphDef.h:112); periodic axes: dp/2 (JPartsLoad4.cpp:389-399)
    double dmin[3],dmax[3],rmin[3],rmax[3];
    for(int k=0;k<3;k++){ const double b=(((p.periactive>>k)&1)? p.dp/2.: border); rmin[k]=cmin[k]-b; rmax[k]=cmax[k]+b; dmin[k]=rmin[k]; dmax[k]=rmax[k]; }
    // re-read <simulationdomain> to apply default+-value / default+-% / fixed
    std::unique_ptr<XmlNode> root=xml_load(s->casepath+".xml");
    const XmlNode* sd=root->path("execution.parameters.simulationdomain");
    const char* ax[3]={"x","y","z"};
    double prcmin[3]={0,0,0},prcmax[3]={0,0,0},valmin[3]={0,0,0},valmax[3]={0,0,0},fixmin[3],fixmax[3]; bool hasfixmin[3]={false,false,false},hasfixmax[3]={false,false,false};
    if(sd)for(int k=0;k<3;k++){
      double v; int m;
      const XmlNode* a=sd->child("posmin"); const XmlNode* b=sd->child("posmax");
      m=parse_posvalue(a? a->attr(ax[k],"default"): "default",true,v);  if(m==1){ fixmin[k]=v; hasfixmin[k]=true; } else if(m==2)valmin[k]=v; else if(m==3)prcmin[k]=v/100;
      m=parse_posvalue(b? b->attr(ax[k],"default"): "default",false,v); if(m==1){ fixmax[k]=v; hasfixmax[k]=true; } else if(m==2)valmax[k]=v; else if(m==3)prcmax[k]=v/100;
    }
    for(int k=0;k<3;k++){
      const double dif=dmax[k]-dmin[k];
      double a=dmin[k]-dif*prcmin[k], b=dmax[k]+dif*prcmax[k];
      a=a-valmin[k]; b=b+valmax[k];
      if(hasfixmin[k])a=fixmin[k];
      if(hasfixmax[k])b=fixmax[k];
      if(a>rmin[k] || b<rmax[k])throw std::runtime_error("Domain limits are not valid.");
      if((p.periactive>>k)&1){ a=rmin[k]; b=rmax[k]; }   // periodic axes keep the particle limits (JSph.cpp:1472-1475)
      p.maprealposmin[k]=a; p.maprealposmax[k]=b;
    }
  }
  s->casedp=data->get_double("Dp",p.dp); s->caseh=data->get_double("H",p.kernelh); s->caseb=data->get_double("B",p.cteb);
  s->caserhop0=data->get_double("Rhop0",p.rhopzero); s->casegamma=data->get_double("Gamma",p.gamma);
  s->casemassb=data->get_double("MassBound",p.massbound); s->casemassf=data->get_double("MassFluid",p.massfluid);
  int r=b2s_config_finalize(&p);
  if(r)throw std::runtime_error(std::string("configuration failed: ")+b2s_error_string(r));
  // slab loading: only the particles of [cuts[rank],cuts[rank+1]) reach the device; capacity is sized for the slab
  unsigned npsel=np;
  std::vector<uint32_t> sel_idp;
  std::vector<float> sel_nor;     // BoundNor of the boundary particles that stay on this rank (slab loading)
  if(slab_nranks>0){
    const uint32_t ncx=p.domcells[0];
    const uint32_t halow=uint32_t(p.scelldiv)*(p.tboundary==B2S_BC_MDBC? 2u: 1u);   // halo planes (slab.cu slab_halo_planes)
    if(uint32_t(slab_nranks)*halow>ncx)throw std::runtime_error("more slabs than the cell planes along x allow (a slab is at least as wide as the halo)");
    const double sc=double(p.scell);
    std::vector<uint32_t> cx(np);
    // cuts balance the WORK (fluid particles weigh more than boundary particles, slab_cuts.h), capacities follow the counts
    std::vector<uint64_t> hist(ncx,0), work(ncx,0);
    for(unsigned i=0;i<np;i++){
      const double d=(pos[size_t(i)*3]-p.domposmin[0])/sc; uint32_t c=(d<0? 0u: uint32_t(d)); if(c>=ncx)c=ncx-1; cx[i]=c; hist[c]++;
      work[c]+=(idp[i]>=p.casenbound? B2S_SLAB_WEIGHT_FLUID: B2S_SLAB_WEIGHT_BOUND);
    }
    s->slab_cuts=b2s_slab_compute_cuts(work,slab_nranks,halow);
    s->slab_maxplane=uint32_t(*std::max_element(hist.begin(),hist.end()));
    const uint32_t ca=s->slab_cuts[slab_rank], cb=s->slab_cuts[slab_rank+1];
    unsigned w=0;
    sel_idp.resize(np);
    if(anor){ sel_nor.reserve(size_t(nbound)*3); for(unsigned i=0;i<nbound;i++)if(cx[i]>=ca && cx[i]<cb)for(int k=0;k<3;k++)sel_nor.push_back(((const float*)anor->ptr())[size_t(i)*3+k]); }
    for(unsigned i=0;i<np;i++)if(cx[i]>=ca && cx[i]<cb){
      sel_idp[w]=idp[i]; code[w]=code[i]; for(int k=0;k<3;k++)pos[size_t(w)*3+k]=pos[size_t(i)*3+k]; for(int k=0;k<4;k++)velrho[size_t(w)*4+k]=velrho[size_t(i)*4+k];
      w++;
    }
    npsel=w; idp=sel_idp.data();
    // room for: ghosts (scelldiv planes each side), migrants, and a 25% drift of the owned count before re-cutting
    extra_capacity+=npsel/4+4u*halow*s->slab_maxplane+4096u;
  }
  if(p.periactive){
    if(p.casenfloat)throw std::runtime_error("periodic boundaries with floating bodies are not supported by this build");
    if(p.tboundary==B2S_BC_MDBC && slab_nranks>1)throw std::runtime_error("periodic boundaries with mDBC and slab decomposition are not supported by this build");
    if(slab_nranks>1 && (p.periactive&1))throw std::runtime_error("periodic boundaries along x (the slab axis) with slab decomposition are not supported by this build");
    // duplicates: the particles within KernelSize of a periodic face, both faces (2*ks/size of the particles per axis), and
    // during a divide the previous set is still in the arrays next to the new one (x2), plus corner copies -> 5*ks/size per axis
    double frac=0;
    for(int k=0;k<3;k++)if((p.periactive>>k)&1)frac+=5.0*double(p.kernelsize)/std::max(p.maprealsize[k],double(p.kernelsize));
    extra_capacity+=unsigned(std::min(3.0,frac)*npsel)+4096u;
  }
  r=b2s_create(&p,device,npsel+extra_capacity,&s->ctx);
  if(r)throw std::runtime_error(std::string("b2s_create failed: ")+b2s_error_string(r));
  if(s->variant>=0 && (r=b2s_set_interaction_variant(s->ctx,s->variant))!=0)throw std::runtime_error("b2s_set_interaction_variant failed");
  // floating bodies (JSph::InitFloatings subset, JSph.cpp:1190-1290); the radii were taken from the full particle set above
  if(p.casenfloat){
    std::vector<b2s_floating> fts;
    size_t kf=0;
    for(auto& b:s->blocks)if(b.type==2){
      b2s_floating f; memset(&f,0,sizeof(f));
      f.begin=b.begin; f.count=b.count; f.mass=float(b.massbody); f.massp=float(b.masspart);
      f.constraints=(b.transfree[0]?0:1)|(b.transfree[1]?0:2)|(b.transfree[2]?0:4)|(b.rotfree[0]?0:8)|(b.rotfree[1]?0:16)|(b.rotfree[2]?0:32);
      for(int k=0;k<3;k++){ f.center[k]=b.center[k]; f.fvel[k]=float(b.linvelini[k]); f.fomega[k]=float(b.angvelini[k]); }
      for(int k=0;k<9;k++)f.inertiaini[k]=float(b.inertia[k]);
      f.radius=ftradius[kf++];
      fts.push_back(f);
    }
    if(s->partbegin){
      // JSph::InitFloatingsRestart (JSph.cpp:2373-2391): centre, linear and angular velocity of PART <begin> and the radius of
      // the head replace the initial values; the constant data must match (JPartFloatInfoBi4Load::CheckHeadData)
      std::string dir=s->partbegindir; if(!dir.empty() && dir[dir.size()-1]!='/')dir+='/';
      const std::string ffile=dir+"PartFloatInfo.ibi4";
      std::vector<std::unique_ptr<Bi4Item>> fl=bi4_load_list(ffile);
      if(fl.empty() || fl[0]->get_uint("FtCount")!=fts.size())throw std::runtime_error("restart: "+ffile+" does not hold the floating bodies of this case");
      const Bi4Array* hb=fl[0]->array("Beginp"); const Bi4Array* hc=fl[0]->array("Countp"); const Bi4Array* hm=fl[0]->array("Mass"); const Bi4Array* hp=fl[0]->array("Massp"); const Bi4Array* hr=fl[0]->array("Radius");
      if(!hb||!hc||!hm||!hp||!hr||hb->count!=fts.size())throw std::runtime_error("restart: head arrays missing in "+ffile);
      const Bi4Item* it=NULL;
      for(size_t k=1;k<fl.size();k++)if(fl[k]->find("Cpart") && fl[k]->get_uint("Cpart")==s->partbegin)it=fl[k].get();
      if(!it)throw std::runtime_error("restart: PART not found in "+ffile);
      const Bi4Array* ac=it->array("center"); const Bi4Array* av=it->array("fvel"); const Bi4Array* ao=it->array("fomega");
      if(!ac||!av||!ao||ac->count!=fts.size())throw std::runtime_error("restart: floating data missing in "+ffile);
      for(size_t cf=0;cf<fts.size();cf++){
        b2s_floating& f=fts[cf];
        if(((const uint32_t*)hb->ptr())[cf]!=f.begin || ((const uint32_t*)hc->ptr())[cf]!=f.count || ((const float*)hm->ptr())[cf]!=f.mass || ((const float*)hp->ptr())[cf]!=f.massp)
          throw std::runtime_error("restart: the constant data of the floating bodies in "+ffile+" do not match the case");
        for(int k=0;k<3;k++){ f.center[k]=((const double*)ac->ptr())[cf*3+k]; f.fvel[k]=((const float*)av->ptr())[cf*3+k]; f.fomega[k]=((const float*)ao->ptr())[cf*3+k]; }
        if(((const float*)hr->ptr())[cf]>0)f.radius=((const float*)hr->ptr())[cf];   // files of the reference hold 0 here
      }
    }
    r=b2s_set_floatings(s->ctx,unsigned(fts.size()),fts.data());
    if(r)throw std::runtime_error(std::string("b2s_set_floatings failed: ")+b2s_last_error(s->ctx));
  }
  for(auto& m:s->motions){
    const PartBlock* blk=NULL;
    for(auto& b:s->blocks)if(b.type==1 && b.refmotion==m.ref)blk=&b;
    if(!blk)continue;
    m.m.idbegin=blk->begin; m.m.count=blk->count;
    r=b2s_add_motion(s->ctx,&m.m);
    if(r)throw std::runtime_error(std::string("b2s_add_motion failed: ")+b2s_last_error(s->ctx));
  }
  for(auto& d:s->dampings){
    if(d.type==0)r=b2s_add_damping_plane(s->ctx,d.a,d.b,d.over,d.redumax,d.factor,d.usedomain? d.domain: NULL);
    else r=b2s_add_damping_box(s->ctx,d.a,d.b,d.c,d.d,d.over,d.redumax,d.factor,d.directions);
    if(r)throw std::runtime_error(std::string("damping zone failed: ")+b2s_last_error(s->ctx));
  }
  for(auto& g:s->gauges){
    if(g.computedt<0)g.computedt=s->timepart;     // JGaugeSystem::ResetCfgDefault
    if(g.computeend<0)g.computeend=s->timemax;
    if(g.outputdt<0)g.outputdt=s->timepart;
    if(g.outputend<0)g.outputend=s->timemax;
    if(g.type==1)r=b2s_gauge_add_velocity(s->ctx,g.point,&g.id);
    else if(g.type==4)r=b2s_gauge_add_maxz(s->ctx,g.point0,g.height,g.distlimit,&g.id);
    else if(g.type==3){
      // block of the target mkbound and the code its particles carry (same numbering as the code array built above)
      const PartBlock* blk=NULL; unsigned idx[4]={0,0,0,0}, bidx=0;
      for(auto& b:s->blocks){ if(b.type!=3 && b.mktype==g.mkbound && !blk){ blk=&b; bidx=idx[b.type]; } idx[b.type]++; }
      if(!blk)throw std::runtime_error("gauge force '"+g.name+"': mkbound is unknown");
      if(blk->type!=0 && blk->type!=1)throw std::runtime_error("gauge force '"+g.name+"': only fixed or moving particles are allowed");
      const uint32_t tcode[4]={0x0000,0x0800,0x1000,0x1800};
      r=b2s_gauge_add_force(s->ctx,blk->begin,blk->count,tcode[blk->type]|bidx,&g.id);
    }
    else r=b2s_gauge_add_swl(s->ctx,g.point0,g.point2,g.pointdp,g.masslimit,&g.id);
    if(r)throw std::runtime_error(std::string("gauge '")+g.name+"' failed: "+b2s_last_error(s->ctx));
  }
  std::vector<float> nor;
  const float* norp=NULL;
  if(p.tboundary==B2S_BC_MDBC){
    if(!anor)throw std::runtime_error("No normal data for mDBC in the input file.");
    if(anor->count!=nbound)throw std::runtime_error("Size of loaded BoundNor data does not match number of boundary particles.");
    if(slab_nranks>0)nor=sel_nor;
    else nor.assign((const float*)anor->ptr(),(const float*)anor->ptr()+size_t(nbound)*3);
    if(!s->partbegin)for(auto& v:nor)v*=2.f;   // boundary->interface normals doubled to reach the ghost node (JSph.cpp:1417)
    else{
      // restart: the normals of that PART, already doubled (moving boundaries have turned theirs), from PartExtra_XXXX.bi4
      // (JSph::ConfigBoundNormals, JSph.cpp:1396-1403; JDsExtraDataLoad::LoadNormals)
      if(slab_nranks>0)throw std::runtime_error("restart with mDBC and slab decomposition is not supported by this build");
      std::string dir=s->partbegindir; if(!dir.empty() && dir[dir.size()-1]!='/')dir+='/';
      char fe[64]; snprintf(fe,sizeof(fe),"PartExtra_%04u.bi4",s->partbegin);
      struct stat st; if(stat((dir+fe).c_str(),&st)!=0){ char m[128]; snprintf(m,sizeof(m),"No extra data available to restart at PART_%04u with mDBC.",s->partbegin); throw std::runtime_error(m); }
      std::unique_ptr<Bi4Item> ex=bi4_load(dir+fe);
      const Bi4Array* qn=ex->array("Normals");
      const uint32_t nsize=nbound-p.casenfloat;
      if(ex->get_uint("CaseNbound")!=nbound || !qn || qn->count<nsize)throw std::runtime_error(std::string("restart: ")+fe+" does not belong to this case");
      std::fill(nor.begin(),nor.end(),0.f);
      // (boundary particles are never excluded while a run goes on, so slot i of the restarted boundary is id i)
      for(unsigned i=0;i<nbound && i<np;i++)if(idp[i]<nsize)for(int k=0;k<3;k++)nor[size_t(i)*3+k]=((const float*)qn->ptr())[size_t(idp[i])*3+k];
    }
    norp=nor.data();
  }
  r=b2s_upload_particles(s->ctx,npsel,idp,code.data(),pos.data(),velrho.data(),norp);
  if(r)throw std::runtime_error(std::string("b2s_upload_particles failed: ")+b2s_last_error(s->ctx));
  s->np0=np;
  if(!s->dirdata.empty()){   // excluded particles -> PartOut_000.obi4 (a slab rank collects what its own particles lose: PartOut_pRR_000.obi4)
    enable_out(s,slab_nranks);
  }
  if(s->partbegin){   // TimeStepIni / SymplecticDtPre / PartIni of the restart (JSph.cpp:2214-2225)
    r=b2s_set_time(s->ctx,(s->partbeginfirst? s->partbegintime: 0.0),s->partbegindtpre);
    if(r)throw std::runtime_error(std::string("b2s_set_time failed: ")+b2s_last_error(s->ctx));
    s->part=s->partbeginfirst;
  }
  if(slab_nranks>0){
    if(slab_nranks>1 && s->timepartextra>=0)throw std::runtime_error("TimeOutExtra / -toutx (PartFloatInfo2, PartMotionRef2) is not written in slab-decomposed runs: remove it or run on one GPU");
    r=b2s_slab_config(s->ctx,slab_rank,slab_nranks,s->slab_cuts[slab_rank],s->slab_cuts[slab_rank+1]);
    if(r)throw std::runtime_error(std::string("b2s_slab_config failed: ")+b2s_last_error(s->ctx));
  }
}

// JDsExtraDataSave::Config / CheckSave (JDsExtraData.cpp:70-88): "<n>" = every n-th PART, otherwise a list "1,4,7-9" (JRangeFilter);
// PART 0 never carries extra data.
static bool svextra_check(const std::string& cfg,unsigned cpart){
  if(cfg.empty() || !cpart)return false;
  if(cfg.find_first_not_of("0123456789")==std::string::npos){ const unsigned n=unsigned(atoi(cfg.c_str())); return n && cpart%n==0; }
  std::stringstream ss(cfg); std::string tok;
  while(std::getline(ss,tok,',')){
    if(tok.empty())continue;
    const size_t d=tok.find('-');
    if(d==std::string::npos){ if(unsigned(atoi(tok.c_str()))==cpart)return true; }
    else{ const unsigned a=unsigned(atoi(tok.substr(0,d).c_str())), b=unsigned(atoi(tok.substr(d+1).c_str())); if(cpart>=a && cpart<=b)return true; }
  }
  return false;
}

// ---- PART output (JSph::SavePartData / JPartDataBi4) ----------------------------------------------------
// One record of the extra-frequency body output (JSphCpuSingle.cpp:1201-1206: PartFloatSave->AddDataExtra + PartMotionSave->AddDataExtraCpu
// at the end of a step that reached TimePartExtraNext): the state of the floating bodies and the positions of the reference particles.
static void extra_collect(b2s_sim* s,const b2s_info& inf){
  const b2s_params& p=s->prm;
  if(p.casenfloat){
    std::vector<const PartBlock*> fb; for(auto& b:s->blocks)if(b.type==2)fb.push_back(&b);
    const unsigned ftc=unsigned(fb.size());
    std::vector<b2s_floating> fts(ftc);
    if(b2s_get_floatings(s->ctx,ftc,fts.data()))throw std::runtime_error(std::string("b2s_get_floatings failed: ")+b2s_last_error(s->ctx));
    b2s_sim::ExtraStore& x=s->xfloat;
    x.time.push_back(inf.timestep); x.step.push_back(inf.nstep);
    for(unsigned cf=0;cf<ftc;cf++)for(int k=0;k<3;k++){
      x.cen.push_back(fts[cf].center[k]);
      const float v[10]={fts[cf].fvel[k],fts[cf].fomega[k],fts[cf].facelin[k],fts[cf].faceang[k],0.f,0.f,fts[cf].fluforcelin[k],fts[cf].fluforceang[k],fts[cf].preacelin[k],fts[cf].preaceang[k]};
      for(int a=0;a<10;a++)x.f3[a].push_back(v[a]);
    }
  }
  if((p.casenmoving || p.casenfloat) && s->motref_ready){
    const std::vector<uint32_t> ids=s->motref.ref_ids();
    std::vector<double> pr(ids.size()*3);
    if(b2s_get_pos_byid(s->ctx,uint32_t(ids.size()),ids.data(),pr.data()))throw std::runtime_error(std::string("b2s_get_pos_byid failed: ")+b2s_last_error(s->ctx));
    for(size_t k=0;k<ids.size();k++)if(pr[k*3]!=pr[k*3])throw std::runtime_error("PartMotionRef2: reference particle "+std::to_string(ids[k])+" is not in the simulation any more");
    b2s_sim::ExtraStore& x=s->xmotion;
    x.time.push_back(inf.timestep); x.step.push_back(inf.nstep); x.posref.insert(x.posref.end(),pr.begin(),pr.end());
  }
}
// JSph::TimeOutExtraUpdate (JSph.cpp:859-870)
static void extra_update_next(b2s_sim* s,double t){
  if(s->timepartextra<0)return;
  double tn=t;
  if(s->timepartextra){ unsigned ct=unsigned(t/s->timepartextra); tn=s->timepartextra*ct; for(;tn<=t;ct++)tn=s->timepartextra*ct; }
  s->timepartextranext=tn;
}

static void save_part(b2s_sim* s){
  const b2s_params& p=s->prm;
  // a slab rank holds 1/N of the case: its buffers follow what it holds (b2s_download_particles_byid returns at most np)
  unsigned cap=p.casenp;
  if(s->npiece>1){ b2s_info i0; if(b2s_get_info(s->ctx,&i0)==0)cap=std::min<unsigned>(p.casenp,i0.np); }
  std::vector<uint32_t> idp(cap); std::vector<double> pos(size_t(cap)*3); std::vector<float> vel(size_t(cap)*3), rho(cap);
  uint32_t n=0;
  int r=b2s_download_particles_byid(s->ctx,&n,idp.data(),pos.data(),vel.data(),rho.data());
  if(r)throw std::runtime_error(std::string("download failed: ")+b2s_last_error(s->ctx));
  b2s_info inf; b2s_get_info(s->ctx,&inf);
  // Item/value order and types follow the reference's files exactly (JPartDataBi4::ConfigBasic/ConfigParticles/AddPartInfo,
  // JPartDataHead::SaveFile, JSph::AddBasicInfo; checked against reference output by tests/test_part_files.py).
  Bi4Item d("JPartDataBi4");
  const double zero3[3]={0,0,0};
  const bool pieces=(s->npiece>1);   // one piece per slab rank: list files that need the whole case are left to single-GPU runs
  d.set_uint("Piece",s->piece); d.set_uint("Npiece",s->npiece); d.set_text("RunCode",s->runcode); d.set_text("Date",s->date);
  d.set_text("AppName","b200sph v0.1 (DualSPHysics v5.4 case interface)"); d.set_text("CaseName",s->casename); d.set_bool("Data2d",p.simulate2d!=0); d.set_double("Data2dPosY",s->posy2d);
  d.set_double3("MapPosMin",p.maprealposmin); d.set_double3("MapPosMax",p.maprealposmax);
  d.set_int("PeriMode",p.periactive); d.set_double3("PeriXinc",(p.periactive&1)? p.perixinc: zero3); d.set_double3("PeriYinc",(p.periactive&2)? p.periyinc: zero3); d.set_double3("PeriZinc",(p.periactive&4)? p.perizinc: zero3);   // TpPeriMode = axis bit mask (DualSphDef.h)
  d.set_int("AxisDiv",0);
  d.set_ullong("CaseNp",p.casenp); d.set_ullong("CaseNfixed",p.casenfixed); d.set_ullong("CaseNmoving",p.casenmoving);
  d.set_ullong("CaseNfloat",p.casenfloat); d.set_ullong("CaseNfluid",p.casenfluid);
  d.set_double3("CasePosMin",s->caseposmin); d.set_double3("CasePosMax",s->caseposmax);
  d.set_bool("NpDynamic",false); d.set_bool("ReuseIds",false);
  d.set_double("Dp",s->casedp); d.set_double("H",s->caseh); d.set_double("B",s->caseb); d.set_double("Rhop0",s->caserhop0);
  d.set_double("Gamma",s->casegamma); d.set_double("MassBound",s->casemassb); d.set_double("MassFluid",s->casemassf);
  d.set_bool("Symmetry",false); d.set_bool("Splitting",false);
  char nm[32]; snprintf(nm,sizeof(nm),"PART_%04u",s->part);
  Bi4Item* pt=d.add_item(nm);
  // particles excluded since the previous PART (JSph::AddParticlesOut / PartsOut): drained from the device buffer
  uint32_t nout=0; unsigned nmot[4]={0,0,0,0};
  std::vector<uint32_t> oid; std::vector<double> opos; std::vector<float> ovel, orho; std::vector<uint8_t> omot;
  {
    const uint32_t ocap=(s->out_cap? s->out_cap: std::max(p.casenfluid,1024u));
    oid.resize(ocap); opos.resize(size_t(ocap)*3); ovel.resize(size_t(ocap)*3); orho.resize(ocap); omot.resize(ocap);
    r=b2s_out_fetch(s->ctx,ocap,&nout,oid.data(),opos.data(),ovel.data(),orho.data(),omot.data());
    if(r)throw std::runtime_error(std::string("b2s_out_fetch failed: ")+b2s_last_error(s->ctx));
    for(uint32_t i=0;i<nout;i++)nmot[omot[i]&3]++;
    s->out_total[0]+=nout; for(int k=1;k<4;k++)s->out_total[k]+=nmot[k];
  }
  pt->set_uint("Cpart",s->part); pt->set_double("TimeStep",inf.timestep); pt->set_uint("Npok",n); pt->set_uint("Nout",nout);
  pt->set_uint("Step",inf.nstep); pt->set_double("RunTime",0);
  pt->set_double3("DomainMin",p.maprealposmin); pt->set_double3("DomainMax",p.maprealposmax);
  pt->set_ullong("NpTotal",p.casenp);
  if(p.tstep==B2S_STEP_SYMPLECTIC)pt->set_double("SymplecticDtPre",inf.symplectic_dtpre);
  // PartInfo statistics of the interval since the previous PART (JSph::AddBasicInfo, JSph.cpp:2937-2975)
  const unsigned dsteps=inf.nstep-s->part_nstep0;
  pt->set_double("dtmean",dsteps? (inf.timestep-s->part_time0)/dsteps: 0.0);
  pt->set_double("dtmin",dsteps? s->part_dtmin: 0.0); pt->set_double("dtmax",dsteps? s->part_dtmax: 0.0);
  pt->set_double("timesim",0);
  pt->set_uint("nct",inf.nct); pt->set_uint("nctsize",inf.nct);
  pt->set_uint("npsim",inf.np); pt->set_uint("npsize",inf.np); pt->set_uint("npnormal",inf.np); pt->set_uint("npsave",n); pt->set_uint("npnew",0);
  pt->set_uint("noutpos",nmot[1]); pt->set_uint("noutrho",nmot[2]); pt->set_uint("noutmov",nmot[3]);
  pt->set_uint("npbin",inf.npbok); pt->set_uint("npbout",inf.npb-inf.npbok); pt->set_uint("npf",inf.np-inf.npb); pt->set_uint("npbper",0); pt->set_uint("npfper",0);
  pt->set_llong("cpualloc",int64_t(cap)*(4+24+12+4));
  s->part_nstep0=inf.nstep; s->part_time0=inf.timestep; s->part_dtmin=DBL_MAX; s->part_dtmax=0;
  std::vector<float> posf;
  pt->add_array_ext("Idp",BT_UINT,n,idp.data());
  if(s->svposdouble)pt->add_array_ext("Posd",BT_DOUBLE3,n,pos.data());
  else{ posf.resize(size_t(n)*3); for(size_t i=0;i<size_t(n)*3;i++)posf[i]=float(pos[i]); pt->add_array_ext("Pos",BT_FLOAT3,n,posf.data()); }
  pt->add_array_ext("Vel",BT_FLOAT3,n,vel.data());
  pt->add_array_ext("Rhop",BT_FLOAT,n,rho.data());
  char fn[64];
  if(pieces)snprintf(fn,sizeof(fn),"Part_p%02u_%04u.bi4",s->piece,s->part); else snprintf(fn,sizeof(fn),"Part_%04u.bi4",s->part);
  bi4_save(s->dirdata+fn,d,"JPartDataBi4");
  const bool firstsave=!s->head_saved;
  const bool fresh=(firstsave && !s->partbeginfirst);   // a restart that continues the numbering appends to the list files
  { // PartInfo.ibi4 (JPartDataBi4::SaveFileInfo, JPartDataBi4.cpp): the head values + one item of values per PART
    Bi4Item hd("JPartDataBi4"); hd.values=d.values;
    Bi4Item pi(nm); pi.hide=true; pi.values=pt->values;
    char fi[64]; if(pieces)snprintf(fi,sizeof(fi),"PartInfo_p%02u.ibi4",s->piece); else snprintf(fi,sizeof(fi),"PartInfo.ibi4");
    bi4_save_listapp(s->dirdata+fi,"JPartDataBi4_Info",hd,&pi,fresh);
  }
  { // PartOut_000.obi4, PartOut_pRR_000.obi4 for the piece of a slab rank (JPartOutBi4Save::GetFileNamePart/ConfigBasic/ConfigParticles/ConfigLimits/SavePartOut, JPartOutBi4Save.cpp:102-232)
    Bi4Item hd("JPartOutBi4");
    hd.set_uint("Piece",s->piece); hd.set_uint("Npiece",s->npiece); hd.set_text("RunCode",s->runcode); hd.set_text("Date",s->date);
    hd.set_text("AppName","b200sph v0.1 (DualSPHysics v5.4 case interface)"); hd.set_bool("Data2d",p.simulate2d!=0);
    hd.set_uint("FmtVersion",180320); hd.set_uint("Block",0);
    hd.set_double3("MapPosMin",p.maprealposmin); hd.set_double3("MapPosMax",p.maprealposmax);
    hd.set_float("RhopMin",p.rhopoutmin); hd.set_float("RhopMax",p.rhopoutmax);
    hd.set_ullong("CaseNp",p.casenp); hd.set_ullong("CaseNfixed",p.casenfixed); hd.set_ullong("CaseNmoving",p.casenmoving);
    hd.set_ullong("CaseNfloat",p.casenfloat); hd.set_ullong("CaseNfluid",p.casenfluid);
    Bi4Item po(nm); po.hide=true;
    po.set_uint("Cpart",s->part); po.set_double("TimeStep",inf.timestep); po.set_uint("Nout",nout);
    std::vector<float> oposf;
    po.add_array_ext("Idp",BT_UINT,nout,oid.data());
    if(s->svposdouble)po.add_array_ext("Posd",BT_DOUBLE3,nout,opos.data());
    else{ oposf.resize(size_t(nout)*3); for(size_t i=0;i<size_t(nout)*3;i++)oposf[i]=float(opos[i]); po.add_array_ext("Pos",BT_FLOAT3,nout,oposf.data()); }
    po.add_array_ext("Vel",BT_FLOAT3,nout,ovel.data()); po.add_array_ext("Rhop",BT_FLOAT,nout,orho.data()); po.add_array_ext("Motive",BT_UCHAR,nout,omot.data());
    char fo[64]; if(pieces)snprintf(fo,sizeof(fo),"PartOut_p%02u_000.obi4",s->piece); else snprintf(fo,sizeof(fo),"PartOut_000.obi4");
    const bool firstout=(pieces? !s->out_saved: firstsave);   // Part_Head.ibi4 (what firstsave follows) comes from rank 0 only
    if(firstout || nout)bi4_save_listapp(s->dirdata+fo,"JPartOutBi4",hd,(nout? &po: NULL),pieces? (firstout && !s->partbeginfirst): fresh);
    s->out_saved=true;
  }
  if(p.casenfloat && (!pieces || s->piece==0)){ // PartFloatInfo.ibi4 (every rank holds the same body state: rank 0 writes it) (JPartFloatInfoBi4Save::Config/MakeBdPartSingle/SaveBdPart, JPartFloatInfoBi4.cpp:464-602)
    std::vector<const PartBlock*> fb; for(auto& b:s->blocks)if(b.type==2)fb.push_back(&b);
    const unsigned ftc=unsigned(fb.size());
    std::vector<b2s_floating> fts(ftc);
    r=b2s_get_floatings(s->ctx,ftc,fts.data());
    if(r)throw std::runtime_error(std::string("b2s_get_floatings failed: ")+b2s_last_error(s->ctx));
    unsigned mkboundfirst=0; for(auto& b:s->blocks)if(b.type!=3){ mkboundfirst=b.mk-b.mktype; break; }
    Bi4Item hd("JPartFloatInfoBi4"), hd2("JPartFloatInfoBi4");   // hd2: the head of PartFloatInfo2.ibi4 (MainFile=false, its own TimeOut)
    for(Bi4Item* h:{&hd,&hd2}){
    Bi4Item& hd=*h;
    hd.set_text("AppName","b200sph v0.1 (DualSPHysics v5.4 case interface)"); hd.set_uint("FormatVer",230729); hd.set_bool("MainFile",h!=&hd2);
    hd.set_double("TimeOut",h!=&hd2? s->timepart: s->timepartextra);
    { Bi4Value v; v.name="MkBoundFirst"; v.type=BT_USHORT; v.i=0; v.u=mkboundfirst; v.d=0; v.d3[0]=v.d3[1]=v.d3[2]=0; hd.values.push_back(v); }
    hd.set_uint("FtCount",ftc); hd.set_uint("FptCount",0);
    }
    std::vector<uint16_t> mkb(ftc); std::vector<uint32_t> beg(ftc),cnt(ftc); std::vector<float> mass(ftc),massp(ftc),radi(ftc);
    std::vector<double> cen(size_t(ftc)*3); std::vector<float> f3[10]; for(auto& v:f3)v.assign(size_t(ftc)*3,0.f);
    for(unsigned cf=0;cf<ftc;cf++){
      mkb[cf]=uint16_t(fb[cf]->mktype); beg[cf]=fts[cf].begin; cnt[cf]=fts[cf].count; mass[cf]=fts[cf].mass; massp[cf]=fts[cf].massp; radi[cf]=fts[cf].radius;
      for(int k=0;k<3;k++){
        cen[cf*3+k]=fts[cf].center[k]; f3[0][cf*3+k]=fts[cf].fvel[k]; f3[1][cf*3+k]=fts[cf].fomega[k]; f3[2][cf*3+k]=fts[cf].facelin[k]; f3[3][cf*3+k]=fts[cf].faceang[k];
        f3[6][cf*3+k]=fts[cf].fluforcelin[k]; f3[7][cf*3+k]=fts[cf].fluforceang[k]; f3[8][cf*3+k]=fts[cf].preacelin[k]; f3[9][cf*3+k]=fts[cf].preaceang[k];
      }
    }
    for(Bi4Item* h:{&hd,&hd2}){
      h->add_array_ext("MkBound",BT_USHORT,ftc,mkb.data()); h->add_array_ext("Beginp",BT_UINT,ftc,beg.data()); h->add_array_ext("Countp",BT_UINT,ftc,cnt.data());
      h->add_array_ext("Mass",BT_FLOAT,ftc,mass.data()); h->add_array_ext("Massp",BT_FLOAT,ftc,massp.data()); h->add_array_ext("Radius",BT_FLOAT,ftc,radi.data());
    }
    Bi4Item pf(nm); pf.hide=true;
    pf.set_uint("Cpart",s->part); pf.set_double("TimeStep",inf.timestep); pf.set_uint("Step",inf.nstep);
    pf.add_array_ext("center",BT_DOUBLE3,ftc,cen.data());
    static const char* fn3[10]={"fvel","fomega","facelin","faceang","extforcelin","extforceang","fluforcelin","fluforceang","preacelin","preaceang"};
    for(int a=0;a<10;a++)pf.add_array_ext(fn3[a],BT_FLOAT3,ftc,f3[a].data());
    bi4_save_listapp(s->dirdata+"PartFloatInfo.ibi4","JPartFloatInfoBi4",hd,&pf,fresh);
    if(s->timepartextra>=0){   // PartFloatInfo2.ibi4: what was collected since the last PART (JDsPartFloatSave::SaveDataExtra, JPartFloatInfoBi4Save::SaveStoredData)
      b2s_sim::ExtraStore& x=s->xfloat;
      if(s->xfloat_fresh && x.time.empty()){   // the first PART of the run is also the first extra record (JSph.cpp:2804-2806)
        x.time.push_back(inf.timestep); x.step.push_back(inf.nstep); x.cen.insert(x.cen.end(),cen.begin(),cen.end());
        for(int a=0;a<10;a++)x.f3[a].insert(x.f3[a].end(),f3[a].begin(),f3[a].end());
      }
      const unsigned cnt2=unsigned(x.time.size());
      if(cnt2){
        Bi4Item pe(nm); pe.hide=true;
        pe.set_uint("Cpart",s->part);
        if(cnt2==1){ pe.set_double("TimeStep",x.time[0]); pe.set_uint("Step",x.step[0]); }
        else{ pe.set_uint("Count",cnt2); pe.add_array_ext("TimeStep",BT_DOUBLE,cnt2,x.time.data()); pe.add_array_ext("Step",BT_UINT,cnt2,x.step.data()); }
        pe.add_array_ext("center",BT_DOUBLE3,ftc*cnt2,x.cen.data());
        for(int a=0;a<10;a++)pe.add_array_ext(fn3[a],BT_FLOAT3,ftc*cnt2,x.f3[a].data());
        bi4_save_listapp(s->dirdata+"PartFloatInfo2.ibi4","JPartFloatInfoBi4",hd2,&pe,s->xfloat_fresh);
        s->xfloat_fresh=false; x.clear();
      }
    }
  }
  if(!pieces && p.tboundary==B2S_BC_MDBC && svextra_check(s->svextraparts,s->part)){
    // PartExtra_XXXX.bi4 (JSphCpuSingle::SaveExtraData, JSphCpuSingle.cpp:1309-1322; JDsExtraDataSave::InitPartData/AddNormals,
    // JDsExtraData.cpp:96-133): the boundary normals by id as the solver holds them (boundary -> ghost node), what a restart needs
    const uint32_t nsize=p.casenp-p.casenfluid-p.casenfloat;   // UseNormalsFt=false: fixed and moving boundary
    std::vector<uint32_t> sid(inf.np); std::vector<float> snor(size_t(inf.np)*4);
    size_t c1=0,c2=0;
    r=b2s_get_array(s->ctx,"idp",sid.data(),sid.size()*4,&c1);
    if(!r)r=b2s_get_array(s->ctx,"boundnor",snor.data(),snor.size()*4,&c2);
    if(r || c1!=c2)throw std::runtime_error(std::string("reading the boundary normals failed: ")+b2s_last_error(s->ctx));
    std::vector<float> vnor(size_t(nsize)*3,0.f);
    for(size_t q=0;q<c1;q++)if(sid[q]<nsize)for(int k=0;k<3;k++)vnor[size_t(sid[q])*3+k]=snor[q*4+k];
    Bi4Item e("JPartExtraBi4");
    e.set_text("AppName","b200sph v0.1 (DualSPHysics v5.4 case interface)"); e.set_uint("FormatVer",211030);
    e.set_uint("CaseNbound",p.casenp-p.casenfluid); e.set_uint("CaseNfloat",p.casenfloat);
    e.set_int("Cpart",int(s->part)); e.set_uint("Step",inf.nstep); e.set_double("TimeStep",inf.timestep); e.set_bool("UseNormalsFt",false);
    e.add_array_ext("Normals",BT_FLOAT3,nsize,vnor.data());
    char fe[64]; snprintf(fe,sizeof(fe),"PartExtra_%04u.bi4",s->part);
    bi4_save(s->dirdata+fe,e,"JPartExtraBi4");
  }
  if(!pieces && (p.casenmoving || p.casenfloat)){ // PartMotionRef.ibi4 (JSph::ConfigSaveData, JSph.cpp:2772-2791; JSphCpuSingle.cpp:1272-1274)
    const uint32_t nb=uint32_t(std::lower_bound(idp.begin(),idp.begin()+n,p.casenp-p.casenfluid)-idp.begin());   // ids ascending: the boundary first
    unsigned mkboundfirst=0; for(auto& b:s->blocks)if(b.type!=3){ mkboundfirst=b.mk-b.mktype; break; }
    if(!s->motref_ready){   // reference particles from the state of the first PART of the run (the initial or the restarted one)
      for(int ty=1;ty<=2;ty++)for(auto& b:s->blocks)if(b.type==ty)s->motref.add_body(ty==2,uint16_t(b.mktype),b.begin,b.count);
      s->motref.config(nb,idp.data(),pos.data());
      s->motref_ready=true;
    }
    s->motref.save_part(s->dirdata,"b200sph v0.1 (DualSPHysics v5.4 case interface)",s->timepart,mkboundfirst,fresh,
                        s->part,inf.timestep,inf.nstep,nb,idp.data(),pos.data());
    if(s->timepartextra>=0){   // PartMotionRef2.ibi4 (JDsPartMotionSave::SaveDataExtra)
      b2s_sim::ExtraStore& x=s->xmotion;
      if(s->xmotion_fresh && x.time.empty()){   // the first PART of the run is also the first extra record (JSph.cpp:2787-2788)
        const std::vector<uint32_t> ids=s->motref.ref_ids();
        x.time.push_back(inf.timestep); x.step.push_back(inf.nstep);
        for(uint32_t id:ids){
          const uint32_t* q=std::lower_bound(idp.data(),idp.data()+nb,id);
          if(q==idp.data()+nb || *q!=id)throw std::runtime_error("PartMotionRef2: reference particle "+std::to_string(id)+" is not in the simulation any more");
          for(int k=0;k<3;k++)x.posref.push_back(pos[size_t(q-idp.data())*3+k]);
        }
      }
      if(!x.time.empty()){
        s->motref.save_extra(s->dirdata,"b200sph v0.1 (DualSPHysics v5.4 case interface)",s->timepartextra,mkboundfirst,s->xmotion_fresh,
                             s->part,unsigned(x.time.size()),x.time.data(),x.step.data(),x.posref.data());
        s->xmotion_fresh=false; x.clear();
      }
    }
  }
  if(!s->head_saved && s->piece!=0)s->head_saved=true;   // Part_Head.ibi4 comes from rank 0
  if(!s->head_saved){
    s->head_saved=true;
    // Part_Head.ibi4 (JPartDataHead::SaveFile, JPartDataHead.cpp): head values + MkBlocks
    Bi4Item h("JPartDataHead");
    h.set_uint("FmtVersion",180324);
    h.set_text("AppName","b200sph v0.1 (DualSPHysics v5.4 case interface)"); h.set_text("Date",s->date); h.set_text("RunCode",s->runcode);
    h.set_text("CaseName",s->casename); h.set_bool("Data2d",p.simulate2d!=0); h.set_double("Data2dPosY",s->posy2d);
    h.set_uint("Npiece",s->npiece); h.set_uint("FirstPart",0);
    h.set_double3("CasePosMin",s->caseposmin); h.set_double3("CasePosMax",s->caseposmax); h.set_bool("NpDynamic",false); h.set_bool("ReuseIds",false);
    h.set_double3("MapPosMin",p.maprealposmin); h.set_double3("MapPosMax",p.maprealposmax);
    h.set_int("PeriMode",p.periactive); h.set_double3("PeriXinc",(p.periactive&1)? p.perixinc: zero3); h.set_double3("PeriYinc",(p.periactive&2)? p.periyinc: zero3); h.set_double3("PeriZinc",(p.periactive&4)? p.perizinc: zero3);
    h.set_uint("ViscoType",unsigned(p.tvisco)); h.set_float("ViscoValue",p.visco); h.set_float("ViscoBoundFactor",p.viscoboundfactor);
    h.set_bool("Symmetry",false); h.set_bool("Splitting",false);
    h.set_double("Dp",s->casedp); h.set_double("H",s->caseh); h.set_double("B",s->caseb); h.set_double("Gamma",s->casegamma); h.set_double("RhopZero",s->caserhop0);
    h.set_double("MassBound",s->casemassb); h.set_double("MassFluid",s->casemassf);
    h.set_float3("Gravity",p.gravity);
    h.set_ullong("CaseNp",p.casenp); h.set_ullong("CaseNfixed",p.casenfixed); h.set_ullong("CaseNmoving",p.casenmoving);
    h.set_ullong("CaseNfloat",p.casenfloat); h.set_ullong("CaseNfluid",p.casenfluid);
    Bi4Item* mk=h.add_item("MkBlocks");
    mk->set_uint("Count",unsigned(s->blocks.size()));
    unsigned c=0;
    for(auto& b:s->blocks){
      char bn[32]; snprintf(bn,sizeof(bn),"MkBlock_%03u",c++);
      Bi4Item* it=mk->add_item(bn);
      static const char* tn[4]={"Fixed","Moving","Floating","Fluid"};
      it->set_text("Type",tn[b.type]); it->set_uint("Mk",b.mk); it->set_uint("MkType",b.mktype); it->set_uint("Count",b.count);
    }
    bi4_save(s->dirdata+"Part_Head.ibi4",h,"JPartDataHead");
  }
}

// ---- gauges: JGaugeItem::Update/SetTimeStep/Output/StoreResult + SaveResults (JDsGaugeItem.h:274-275, .cpp:239-249,360-403,654-700)
static void gauges_record(b2s_sim* s,double t){
  for(auto& g:s->gauges){
    if(!(t>=g.computenext && g.computestart<=t && t<=g.computeend))continue;
    if(g.computedt){ const unsigned nt=unsigned(t/g.computedt); g.computenext=g.computedt*nt; if(g.computenext<=t)g.computenext=g.computedt*(nt+1); }
    float res[3]={0,0,0};
    if(b2s_gauge_result(s->ctx,g.id,res))throw std::runtime_error("gauge result not available");
    if(!(g.output && t>=g.outputnext && g.outputstart<=t && t<=g.outputend) || s->dirout.empty())continue;
    if(s->piece!=0){   // every slab rank holds the same results: rank 0 writes the files, the others only keep the schedule
      if(g.outputdt){ const unsigned nt=unsigned(t/g.outputdt); g.outputnext=g.outputdt*nt; if(g.outputnext<=t)g.outputnext=g.outputdt*(nt+1); }
      continue;
    }
    const std::string fn=s->dirout+(g.type==1? "GaugesVel_": (g.type==3? "GaugesForce_": (g.type==4? "GaugesMaxZ_": "GaugesSWL_")))+g.name+".csv";
    FILE* f=fopen(fn.c_str(),g.headdone? "a": "w");
    if(!f)throw std::runtime_error("cannot write "+fn);
    if(!g.headdone){
      fputs(g.type==1? "time [s];velx [m/s];vely [m/s];velz [m/s];posx [m];posy [m];posz [m]\n"
            : g.type==3? "time [s];force [N];forcex [N];forcey [N];forcez [N]\n"
            : g.type==4? "time [s];zmax [m];posx [m];posy [m];posz [m]\n"
                     : "time [s];swlx [m];swly [m];swlz [m];pos0x [m];pos0y [m];pos0z [m];pos2x [m];pos2y [m];pos2z [m]\n",f);
      g.headdone=true;
    }
    if(g.type==1)fprintf(f,"%f;%g;%g;%g;%g;%g;%g\n",t,res[0],res[1],res[2],float(g.point[0]),float(g.point[1]),float(g.point[2]));
    else if(g.type==4)fprintf(f,"%f;%g;%g;%g;%g\n",t,res[0],float(g.point0[0]),float(g.point0[1]),float(g.point0[2]));
    else if(g.type==3)fprintf(f,"%f;%g;%g;%g;%g\n",t,sqrtf(res[0]*res[0]+res[1]*res[1]+res[2]*res[2]),res[0],res[1],res[2]);
    else fprintf(f,"%f;%g;%g;%g;%g;%g;%g;%g;%g;%g\n",t,res[0],res[1],res[2],float(g.point0[0]),float(g.point0[1]),float(g.point0[2]),float(g.point2[0]),float(g.point2[1]),float(g.point2[2]));
    fclose(f);
    if(g.outputdt){ const unsigned nt=unsigned(t/g.outputdt); g.outputnext=g.outputdt*nt; if(g.outputnext<=t)g.outputnext=g.outputdt*(nt+1); }
  }
}

// ---- C entry points ---------------------------------------------------------------------------------------
extern "C" const char* b2s_host_last_error(void){ return g_host_err.c_str(); }

extern "C" int b2s_sim_load(const char* casepath,const char* dirout,const char* options,int device,uint32_t extra_capacity,b2s_sim** out){
  if(!casepath||!out)return B2S_ERR_ARG;
  *out=NULL;
  b2s_sim* s=new b2s_sim();
  try{
    s->casepath=casepath;
    if(s->casepath.size()>4 && lower(s->casepath.substr(s->casepath.size()-4))==".xml")s->casepath.resize(s->casepath.size()-4);
    size_t sl=s->casepath.find_last_of('/');
    s->casename=(sl==std::string::npos? s->casepath: s->casepath.substr(sl+1));
    s->dirout=(dirout&&dirout[0]? dirout: "");
    if(!s->dirout.empty()){ if(s->dirout[s->dirout.size()-1]!='/')s->dirout+='/'; s->dirdata=s->dirout+"data/"; mkdirs(s->dirdata); }
    load_case(s,options? options: "");
    load_particles_and_create(s,device,extra_capacity);
    s->timepartnext=0;
  }
  catch(const std::exception& e){ g_host_err=e.what(); if(s->ctx)b2s_destroy(s->ctx); delete s; return B2S_ERR_IO; }
  *out=s;
  return B2S_OK;
}
// Slab loading for one rank of a multi-GPU run: every rank reads the same case, derives the same cuts from the particle
// histogram along x, and uploads only its own slab (device capacity sized for the slab, not for the whole case).
extern "C" int b2s_sim_load_slab(const char* casepath,const char* options,int device,int rank,int nranks,b2s_sim** out){
  if(!casepath||!out||nranks<1||rank<0||rank>=nranks)return B2S_ERR_ARG;
  *out=NULL;
  b2s_sim* s=new b2s_sim();
  try{
    s->casepath=casepath;
    if(s->casepath.size()>4 && lower(s->casepath.substr(s->casepath.size()-4))==".xml")s->casepath.resize(s->casepath.size()-4);
    size_t sl=s->casepath.find_last_of('/');
    s->casename=(sl==std::string::npos? s->casepath: s->casepath.substr(sl+1));
    s->piece=unsigned(rank); s->npiece=unsigned(nranks);
    load_case(s,options? options: "");
    load_particles_and_create(s,device,0,rank,nranks);
  }
  catch(const std::exception& e){ g_host_err=e.what(); if(s->ctx)b2s_destroy(s->ctx); delete s; return B2S_ERR_IO; }
  *out=s;
  return B2S_OK;
}
// Output directory of a slab rank: every rank writes its own piece Part_pRR_XXXX.bi4 / PartInfo_pRR.ibi4, rank 0 also Part_Head.ibi4.
extern "C" int b2s_sim_set_dirout(b2s_sim* s,const char* dirout){
  if(!s||!dirout||!dirout[0])return B2S_ERR_ARG;
  try{ s->dirout=dirout; if(s->dirout[s->dirout.size()-1]!='/')s->dirout+='/'; s->dirdata=s->dirout+"data/"; mkdirs(s->dirdata); enable_out(s,int(s->npiece)); }
  catch(const std::exception& e){ g_host_err=e.what(); return B2S_ERR_IO; }
  return B2S_OK;
}
// Transport of a slab rank (csrc/slab_comm.cu): NCCL communicator from the id rank 0 made, then the first halo exchange.
extern "C" int b2s_sim_slab_connect(b2s_sim* s,const uint8_t* nccl_id128){
  if(!s||!nccl_id128||s->slab_cuts.empty())return B2S_ERR_ARG;
  int r=b2s_slab_comm_init(s->ctx,nccl_id128,s->slab_maxplane);
  if(!r)r=b2s_slab_exchange(s->ctx);
  if(r)g_host_err=b2s_last_error(s->ctx);
  return r;
}
extern "C" int b2s_sim_slab_info(b2s_sim* s,uint32_t* cuts,uint32_t ncuts,uint32_t* max_plane_particles){
  if(!s||s->slab_cuts.empty())return B2S_ERR_STATE;
  if(cuts){ if(ncuts!=s->slab_cuts.size())return B2S_ERR_ARG; for(uint32_t i=0;i<ncuts;i++)cuts[i]=s->slab_cuts[i]; }
  if(max_plane_particles)*max_plane_particles=s->slab_maxplane;
  return B2S_OK;
}
extern "C" int b2s_sim_free(b2s_sim* s){ if(!s)return B2S_OK; if(s->ctx)b2s_destroy(s->ctx); delete s; return B2S_OK; }
extern "C" int b2s_sim_ctx(b2s_sim* s,b2s_ctx** out){ if(!s||!out)return B2S_ERR_ARG; *out=s->ctx; return B2S_OK; }
extern "C" int b2s_sim_params(b2s_sim* s,b2s_params* out){ if(!s||!out)return B2S_ERR_ARG; *out=s->prm; return B2S_OK; }
extern "C" int b2s_sim_stats(b2s_sim* s,uint32_t* parts,uint32_t* excluded4){
  if(!s)return B2S_ERR_ARG;
  if(parts)*parts=s->part;
  if(excluded4)for(int k=0;k<4;k++)excluded4[k]=s->out_total[k];
  return B2S_OK;
}
// Test entry point of the XML reader (include/b2s_host.h): elements and attributes of a file, counted over the whole tree.
static void xml_count(const XmlNode* n,uint32_t& ne,uint32_t& na){ ne++; na+=uint32_t(n->attrs.size()); for(auto& c:n->children)xml_count(c.get(),ne,na); }
extern "C" int b2s_host_xml_stats(const char* file,uint32_t* elements,uint32_t* attributes){
  if(!file)return B2S_ERR_ARG;
  try{ std::unique_ptr<XmlNode> root=xml_load(file); uint32_t ne=0,na=0; xml_count(root.get(),ne,na); if(elements)*elements=ne; if(attributes)*attributes=na; }
  catch(const std::exception& e){ g_host_err=e.what(); return B2S_ERR_IO; }
  return B2S_OK;
}
// Test entry point of the PartMotionRef.ibi4 writer on host arrays (include/b2s_host.h).
extern "C" int b2s_host_motionref_write(const char* dirdata,uint32_t nbodies,const uint8_t* floating,const uint16_t* mkbound,
    const uint32_t* begin,const uint32_t* count,double timeout,uint32_t mkboundfirst,uint32_t nparts,uint32_t n,const uint32_t* idp,
    const double* pos,const double* timestep,const uint32_t* step)
{
  if(!dirdata || !nbodies || !floating || !mkbound || !begin || !count || !nparts || !idp || !pos || !timestep || !step){ g_host_err="b2s_host_motionref_write: null argument"; return B2S_ERR_ARG; }
  try{
    MotionRefSave m;
    for(uint32_t c=0;c<nbodies;c++)m.add_body(floating[c]!=0,mkbound[c],begin[c],count[c]);
    m.config(n,idp,pos);
    for(uint32_t k=0;k<nparts;k++)m.save_part(dirdata,"b200sph v0.1 (DualSPHysics v5.4 case interface)",timeout,mkboundfirst,k==0,k,timestep[k],step[k],n,idp,pos+size_t(k)*n*3);
  }catch(const std::exception& e){ g_host_err=e.what(); return B2S_ERR_IO; }
  return 0;
}

extern "C" int b2s_sim_save_part(b2s_sim* s){
  if(!s)return B2S_ERR_ARG;
  if(s->dirdata.empty()){ g_host_err="no output directory configured"; return B2S_ERR_STATE; }
  try{ save_part(s); s->part++; }catch(const std::exception& e){ g_host_err=e.what(); return B2S_ERR_IO; }
  return B2S_OK;
}

// JSphGpuSingle::Run main loop (JSphGpuSingle.cpp:998-1041): step, PART output at TimeOut, stop at TimeMax / nsteps.
extern "C" int b2s_sim_run(b2s_sim* s,uint32_t nsteps_break,double* time_out,uint32_t* steps_out){
  if(!s)return B2S_ERR_ARG;
  const unsigned nbreak=(nsteps_break? nsteps_break: s->nstepsbreak);
  const bool save=(s->svdata && !s->dirdata.empty());
  try{
    b2s_info inf; int r=b2s_get_info(s->ctx,&inf); if(r){ g_host_err=b2s_last_error(s->ctx); return r; }
    if(!s->gauges.empty() && inf.nstep==0){   // initial measurement (JSphCpuSingle.cpp:1110-1111)
      r=(s->npiece>1? b2s_slab_gauges_compute(s->ctx): b2s_gauges_compute(s->ctx)); if(r){ g_host_err=b2s_last_error(s->ctx); return r; }   // slab ranks: collectively, every rank ends with the same results
      gauges_record(s,inf.timestep);
    }
    if(save && !s->initial_saved){   // the initial PART (PartIni: 0, or the first PART number of a restart)
      save_part(s); s->part++; s->initial_saved=true;
      const double t0=inf.timestep;
      extra_update_next(s,t0);
      if(s->svallsteps)s->timepartnext=t0;
      else{ unsigned ct=unsigned(t0/s->timepart); double tn=s->timepart*ct; for(;tn<=t0;ct++)tn=s->timepart*ct; s->timepartnext=tn; }
    }
    unsigned nstep=0;
    double t=inf.timestep;
    while(t<s->timemax){
      double dt=0;
      r=(s->npiece>1? b2s_slab_step(s->ctx,&dt): b2s_compute_step(s->ctx,&dt));   // slab rank: collective step over NCCL (b2s_sim_slab_connect)
      if(r){ g_host_err=b2s_last_error(s->ctx); return r; }
      r=b2s_get_info(s->ctx,&inf); if(r){ g_host_err=b2s_last_error(s->ctx); return r; }
      t=inf.timestep; nstep++;
      s->part_dtmin=std::min(s->part_dtmin,dt); s->part_dtmax=std::max(s->part_dtmax,dt);
      if(!s->gauges.empty())gauges_record(s,t);
      if(save && s->timepartextra>=0 && t>=s->timepartextranext){ extra_collect(s,inf); extra_update_next(s,t); }
      if(save && t>=s->timepartnext){
        save_part(s); s->part++;
        if(s->svallsteps)s->timepartnext=t;
        else{ unsigned ct=unsigned(t/s->timepart); double tn=s->timepart*ct; for(;tn<=t;ct++)tn=s->timepart*ct; s->timepartnext=tn; }  // JDsOutputTime::GetNextTime
      }
      if(s->npiece<=1 && inf.np<=inf.npb){ g_host_err="no fluid particles left"; break; }
      if(nbreak && nstep>=nbreak)break;
    }
    if(time_out)*time_out=t;
    if(steps_out)*steps_out=nstep;
  }
  catch(const std::exception& e){ g_host_err=e.what(); return B2S_ERR_IO; }
  return B2S_OK;
}
