// ref_probe.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle, "real reference" leg).
//
// Drives the UNMODIFIED DualSPHysics v5.4 CPU solver (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/obj/*.o; no reference source lives in this repository)
// through the same call sequence as JSphCpuSingle::Run / ComputeStep_Sym / ComputeStep_Ver
// (src/source/JSphCpuSingle.cpp:1144-1236, 871-925) and dumps the per-phase state the
// reference itself never writes: cell tables after JCellDivCpuSingle::Divide, per-particle
// ace / ar after Interaction_Forces, the three dt maxima, and the particle state after
// each step.  The dumps are the fixtures the CUDA path and the C restatement are checked
// against (tests/, tests/golden/).
//
// Access to the solver's protected members uses the usual test-harness trick of
// re-declaring access specifiers while including the reference headers; object layout
// is unchanged, and nothing of the reference's behaviour is altered.
//
// Usage: dsph_ref_probe <case path without extension> <dirout> <nsteps> [reference CLI options...]
//   writes <dirout>/probe/<tag>.b2sd  (format: see dump_* below; read by tests/b2sd.py)
#include <string>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <fstream>
#include <vector>
#include <algorithm>
#include <cfloat>
#include <sys/stat.h>

#define protected public
#define private public
#include "JAppInfo.h"
#include "JLog2.h"
#include "JException.h"
#include "JSphCfgRun.h"
#include "JSphCpuSingle.h"
#include "JCellDivCpuSingle.h"
#include "JArraysCpu.h"
#include "JDsTimersCpu.h"
#include "JDsOutputTime.h"
#include "JTimer.h"
#undef protected
#undef private

using namespace std;

JAppInfo AppInfo("DualSPHysics5","v5.4.355","08-04-2025");  // same global the reference's main.cpp defines

// ---------------------------------------------------------------------------------------------
// B2SD dump container: "B2SDUMP1", u32 nentries, then per entry
//   char name[24]; u32 dtype(0=u8,1=u16,2=u32,3=i32,4=f32,5=f64); u64 count; raw data
// ---------------------------------------------------------------------------------------------
struct Dump{
  struct Ent{ string name; unsigned dtype; unsigned long long count; const void* ptr; vector<char> own; };
  vector<Ent> ents;
  static size_t esize(unsigned t){ static const size_t s[6]={1,2,4,4,4,8}; return s[t]; }
  void add(const string& n,unsigned t,unsigned long long c,const void* p){ Ent e; e.name=n; e.dtype=t; e.count=c; e.ptr=p; ents.push_back(e); }
  template<class T> void scalar(const string& n,unsigned t,T v){ Ent e; e.name=n; e.dtype=t; e.count=1; e.own.resize(sizeof(T)); memcpy(e.own.data(),&v,sizeof(T)); e.ptr=NULL; ents.push_back(e); }
  void u32(const string& n,unsigned v){ scalar<unsigned>(n,2,v); }
  void f32(const string& n,float v){ scalar<float>(n,4,v); }
  void f64(const string& n,double v){ scalar<double>(n,5,v); }
  void save(const string& file)const{
    FILE* f=fopen(file.c_str(),"wb"); if(!f){ fprintf(stderr,"probe: cannot write %s\n",file.c_str()); exit(2); }
    fwrite("B2SDUMP1",1,8,f); unsigned n=unsigned(ents.size()); fwrite(&n,4,1,f);
    for(size_t c=0;c<ents.size();c++){
      const Ent& e=ents[c]; char nm[24]; memset(nm,0,24); strncpy(nm,e.name.c_str(),23);
      fwrite(nm,1,24,f); fwrite(&e.dtype,4,1,f); fwrite(&e.count,8,1,f);
      const void* p=(e.ptr? e.ptr: (const void*)e.own.data());
      if(e.count)fwrite(p,esize(e.dtype),size_t(e.count),f);
    }
    fclose(f);
  }
};

// ---------------------------------------------------------------------------------------------
class Probe : public JSphCpuSingle{
public:
  string DirProbe;
  int DumpLevel;   // 0: dt/maxima only, 1: + ace/ar + state, 2: + cell tables
  vector<double> FtDump; vector<float> FtDumpF;

  void DumpCellDiv(Dump& d){
    JCellDivCpuSingle* cd=CellDivSingle;
    d.u32("ncx",cd->Ncx); d.u32("ncy",cd->Ncy); d.u32("ncz",cd->Ncz); d.u32("nct",cd->Nct);
    d.u32("cellmin_x",cd->CellDomainMin.x); d.u32("cellmin_y",cd->CellDomainMin.y); d.u32("cellmin_z",cd->CellDomainMin.z);
    d.u32("cellmax_x",cd->CellDomainMax.x); d.u32("cellmax_y",cd->CellDomainMax.y); d.u32("cellmax_z",cd->CellDomainMax.z);
    d.u32("npb_ignore",cd->NpbIgnore); d.u32("np_final",cd->NpFinal); d.u32("npb_final",cd->NpbFinal);
    d.u32("npf_out",cd->NpfOut); d.u32("divide_full",cd->DivideFull? 1: 0);
    d.u32("domcellcode",cd->DomCellCode);
    if(DumpLevel>=2){
      d.add("cellpart",2,cd->Nptot,cd->CellPart);
      d.add("sortpart",2,cd->Nptot,cd->SortPart);
      d.add("begincell",2,cd->Nctt,cd->BeginCell);
    }
  }
  void DumpState(Dump& d){
    d.u32("np",Np); d.u32("npb",Npb); d.u32("npbok",NpbOk); d.u32("nstep",Nstep);
    d.f64("timestep",TimeStep); d.f64("symplectic_dtpre",SymplecticDtPre);
    if(FtCount){
      // floating bodies: center (f64 x3), fvel, fomega, angles (f32 x3) per body, flattened
      FtDump.assign(size_t(FtCount)*3,0); FtDumpF.assign(size_t(FtCount)*9,0);
      for(unsigned cf=0;cf<FtCount;cf++){
        FtDump[cf*3]=FtObjs[cf].center.x; FtDump[cf*3+1]=FtObjs[cf].center.y; FtDump[cf*3+2]=FtObjs[cf].center.z;
        FtDumpF[cf*9+0]=FtObjs[cf].fvel.x;   FtDumpF[cf*9+1]=FtObjs[cf].fvel.y;   FtDumpF[cf*9+2]=FtObjs[cf].fvel.z;
        FtDumpF[cf*9+3]=FtObjs[cf].fomega.x; FtDumpF[cf*9+4]=FtObjs[cf].fomega.y; FtDumpF[cf*9+5]=FtObjs[cf].fomega.z;
        FtDumpF[cf*9+6]=FtObjs[cf].angles.x; FtDumpF[cf*9+7]=FtObjs[cf].angles.y; FtDumpF[cf*9+8]=FtObjs[cf].angles.z;
      }
      d.add("ft_center",5,FtDump.size(),FtDump.data());
      d.add("ft_velomegaang",4,FtDumpF.size(),FtDumpF.data());
    }
    if(DumpLevel>=1){
      d.add("idp",2,Np,Idp_c->cptr());
      d.add("code",1,Np,Code_c->cptr());
      d.add("dcell",2,Np,Dcell_c->cptr());
      d.add("pos",5,(unsigned long long)Np*3,Pos_c->cptr());
      d.add("velrho",4,(unsigned long long)Np*4,Velrho_c->cptr());
      if(TStep==STEP_Verlet && VelrhoM1_c)d.add("velrhom1",4,(unsigned long long)Np*4,VelrhoM1_c->cptr());
      if(SpsTauRho2_c && TVisco==VISCO_LaminarSPS)d.add("spstau",4,(unsigned long long)Np*6,SpsTauRho2_c->cptr());
      if(BoundNor_c && UseNormals)d.add("boundnor",4,(unsigned long long)Np*3,BoundNor_c->cptr());
    }
  }
  void DumpForces(Dump& d){
    d.u32("np",Np); d.u32("npb",Npb); d.u32("npbok",NpbOk);
    d.f64("velmax",VelMax); d.f64("acemax",AceMax); d.f32("viscdtmax",ViscDtMax);
    if(DumpLevel>=1){
      d.add("idp",2,Np,Idp_c->cptr());
      d.add("ace",4,(unsigned long long)Np*3,Ace_c->cptr());
      d.add("ar",4,Np,Ar_c->cptr());
      d.add("press",4,Np,Press_c->cptr());
      d.add("velrho",4,(unsigned long long)Np*4,Velrho_c->cptr());   // includes mDBC-corrected boundary rho
      if(Sps2Strain_c && TVisco==VISCO_LaminarSPS){
        d.add("spstau_new",4,(unsigned long long)Np*6,SpsTauRho2_c->cptr());
      }
    }
  }
  void Save(const Dump& d,const char* tag,unsigned step){
    char nm[256]; snprintf(nm,sizeof(nm),"%s/s%04u_%s.b2sd",DirProbe.c_str(),step,tag);
    d.save(nm);
  }

  // One Symplectic step: the call order of JSphCpuSingle::ComputeStep_Sym (JSphCpuSingle.cpp:897-932).
  double StepSym(){
    const double dt=SymplecticDtPre;
    if(CaseNmoving)CalcMotion(dt);
    InterStep=INTERSTEP_SymPredictor;
    DemDtForce=dt*0.5f;
    MdbcBoundCorrection(InterStep);
    PreInteraction_Forces(InterStep);
    PreLoopProcedure(InterStep);
    Interaction_Forces(InterStep);
    const double dt_p=DtVariable(false);
    { Dump d; DumpForces(d); d.f64("dt",dt); d.f64("dt_p",dt_p); Save(d,"pred",Nstep); }
    if(Shifting)RunShifting(dt*.5);
    ComputeSymplecticPre(dt);
    if(CaseNfloat)RunFloating(dt*.5,true);
    PosInteraction_Forces();
    InterStep=INTERSTEP_SymCorrector;
    DemDtForce=dt;
    RunCellDivide(true);
    { Dump d; DumpCellDiv(d); DumpState(d); Save(d,"mid",Nstep); }
    MdbcBoundCorrection(InterStep);
    PreInteraction_Forces(InterStep);
    PreLoopProcedure(InterStep);
    Interaction_Forces(InterStep);
    const double dt_c=DtVariable(true);
    { Dump d; DumpForces(d); d.f64("dt",dt); d.f64("dt_c",dt_c); Save(d,"corr",Nstep); }
    if(Shifting)RunShifting(dt);
    ComputeSymplecticCorr(dt);
    if(CaseNfloat)RunFloating(dt,false);
    PosInteraction_Forces();
    if(Damping)RunDamping(dt);
    SymplecticDtPre=min(dt_p,dt_c);
    return(dt);
  }
  // One Verlet step: the call order of JSphCpuSingle::ComputeStep_Ver (JSphCpuSingle.cpp:871-886).
  double StepVer(){
    InterStep=INTERSTEP_Verlet;
    MdbcBoundCorrection(InterStep);
    PreInteraction_Forces(InterStep);
    Interaction_Forces(InterStep);
    const double dt=DtVariable(true);
    { Dump d; DumpForces(d); d.f64("dt",dt); Save(d,"pred",Nstep); }
    if(CaseNmoving)CalcMotion(dt);
    DemDtForce=dt;
    if(Shifting)RunShifting(dt);
    ComputeVerlet(dt);
    if(CaseNfloat)RunFloating(dt,false);
    PosInteraction_Forces();
    if(Damping)RunDamping(dt);
    return(dt);
  }

  // Same initialisation order as JSphCpuSingle::Run (JSphCpuSingle.cpp:1144-1175), then nsteps steps.
  void RunProbe(std::string appname,const JSphCfgRun* cfg,JLog2* log,unsigned nsteps,int dumplevel){
    AppName=appname; Log=log; CfgRun=cfg; DumpLevel=dumplevel;
    Arrays_Cpu=new JArraysCpu(Log);
    Timersc->Config(cfg->SvTimers);
    Timersc->TmStart(TMC_Init);
    LoadConfig(cfg);
    LoadCaseParticles();
    VisuConfig();
    ConfigDomain();
    ConfigRunMode();
    VisuParticleSummary();
    InitRunCpu();
    RunFirstGaugeSystem(TimeStep);
    FreePartsInit();
    UpdateMaxValues();
    SaveData();
    Arrays_Cpu->FreeUnusedArrays();
    Timersc->ResetTimes();
    Timersc->TmStop(TMC_Init);
    PartNstep=0; Part++;
    DirProbe=DirOut+"probe";
    mkdir(DirProbe.c_str(),0777);
    {
      Dump d; DumpCellDiv(d); DumpState(d);
      // constants a reader needs to interpret the dumps
      d.f32("kernelh",KernelH); d.f32("kernelsize",KernelSize); d.f32("scell",Scell); d.u32("scelldiv",unsigned(ScellDiv));
      d.f32("cteb",CteB); d.f32("gamma",Gamma); d.f32("rhopzero",RhopZero); d.f64("cs0",Cs0); d.f32("eta2",Eta2);
      d.f32("massfluid",MassFluid); d.f32("massbound",MassBound); d.f64("dp",Dp); d.f64("cflnumber",CFLnumber);
      d.f32("ddtkh",DDTkh); d.f32("ddtgz",DDTgz); d.f32("visco",Visco); d.f32("viscoboundfactor",ViscoBoundFactor);
      d.f64("dtini",DtIni); d.f64("dtmin",DtMin); d.f32("movlimit",MovLimit);
      d.f32("awen",KWend.awen); d.f32("bwenh",KWend.bwenh);
      d.f64("mapposmin_x",Map_PosMin.x); d.f64("mapposmin_y",Map_PosMin.y); d.f64("mapposmin_z",Map_PosMin.z);
      d.f64("mapposmax_x",Map_PosMax.x); d.f64("mapposmax_y",Map_PosMax.y); d.f64("mapposmax_z",Map_PosMax.z);
      d.f64("maprealmin_x",MapRealPosMin.x); d.f64("maprealmin_y",MapRealPosMin.y); d.f64("maprealmin_z",MapRealPosMin.z);
      d.f64("maprealmax_x",MapRealPosMax.x); d.f64("maprealmax_y",MapRealPosMax.y); d.f64("maprealmax_z",MapRealPosMax.z);
      d.u32("mapcells_x",Map_Cells.x); d.u32("mapcells_y",Map_Cells.y); d.u32("mapcells_z",Map_Cells.z);
      d.u32("tstep",unsigned(TStep)); d.u32("tkernel",unsigned(TKernel)); d.u32("tvisco",unsigned(TVisco));
      d.u32("tdensity",unsigned(TDensity)); d.u32("tboundary",unsigned(TBoundary)); d.u32("slipmode",unsigned(SlipMode));
      d.u32("casenp",CaseNp); d.u32("casenfixed",CaseNfixed); d.u32("casenmoving",CaseNmoving);
      d.u32("casenfloat",CaseNfloat); d.u32("casenfluid",CaseNfluid);
      Save(d,"init",0);
    }
    TimerSim.Start();
    TimerPart.Start();
    PrintHeadPart();
    double tsteps=0;
    for(unsigned s=0;s<nsteps && TimeStep<TimeMax;s++){
      InterStep=(TStep==STEP_Symplectic? INTERSTEP_SymPredictor: INTERSTEP_Verlet);
      JTimer tm; tm.Start();
      const double stepdt=(TStep==STEP_Verlet? StepVer(): StepSym());
      RunGaugeSystem(TimeStep+stepdt);
      if(CaseNmoving)RunMotion(stepdt);
      RunCellDivide(true);
      TimeStep+=stepdt;
      LastDt=stepdt;
      Nstep++;
      tm.Stop(); tsteps+=tm.GetSecs();
      { Dump d; DumpCellDiv(d); DumpState(d); d.f64("stepdt",stepdt); Save(d,"end",Nstep-1); }
      if(TimeStep>=TimePartNext){
        SaveData();
        Part++;
        PartNstep=Nstep;
        TimeStepM1=TimeStep;
        TimePartNext=(SvAllSteps? TimeStep: OutputTime->GetNextTime(TimeStep));
        TimerPart.Start();
      }
      UpdateMaxValues();
    }
    TimerSim.Stop(); TimerTot.Stop();
    printf("PROBE_RESULT steps=%u np=%u seconds=%.6f omp_threads=%d\n",Nstep,Np,tsteps,OmpThreads);
    FinishRun(false);
  }
};

int main(int argc,char** argv){
  if(argc<4){
    printf("usage: %s <case> <dirout> <nsteps> [-dumplevel:N] [reference options]\n",argv[0]);
    return(1);
  }
  const unsigned nsteps=unsigned(atoi(argv[3]));
  int dumplevel=2;
  // rebuild argv for the reference's own option parser: prog case dirout -cpu opts...
  vector<string> args;
  args.push_back(argv[0]); args.push_back(argv[1]); args.push_back(argv[2]); args.push_back("-cpu");
  for(int c=4;c<argc;c++){
    string a=argv[c];
    if(a.compare(0,11,"-dumplevel:")==0)dumplevel=atoi(a.c_str()+11);
    else args.push_back(a);
  }
  vector<char*> av; for(size_t c=0;c<args.size();c++)av.push_back((char*)args[c].c_str());
  int errcode=1;
  AppInfo.ConfigRunPaths(argv[0]);
  const std::string appname=AppInfo.GetFullName();
  JLog2* log=NULL;
  JSphCfgRun cfg;
  cfg.SetFeatureList(JSph::GetFeatureList());
  try{
    cfg.LoadArgv(int(av.size()),av.data());
    AppInfo.ConfigOutput(cfg.CreateDirs,cfg.CsvSepComa,cfg.DirOut,cfg.DirDataOut);
    AppInfo.LogInit(AppInfo.GetDirOut()+"/Run.out");
    log=AppInfo.LogPtr();
    cfg.Cpu=true;
    Probe sph;
    sph.RunProbe(appname,&cfg,log,nsteps,dumplevel);
    errcode=0;
  }
  catch(const char* cad){ fprintf(stderr,"\n*** Exception(chr): %s\n",cad); }
  catch(const string& e){ fprintf(stderr,"\n*** Exception(str): %s\n",e.c_str()); }
  catch(const JException& e){ fprintf(stderr,"\n*** JException: %s\n",e.what()); }
  catch(const exception& e){ fprintf(stderr,"\n*** Exception(exc): %s\n",e.what()); }
  catch(...){ fprintf(stderr,"\n*** Unknown exception\n"); }
  printf("\nFinished execution (code=%d).\n",errcode);
  return(errcode);
}
