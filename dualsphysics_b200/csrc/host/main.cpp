// main.cpp -- b200sph command line: b200sph <case> <dirout> [options]   (same shape as the reference's
// "DualSPHysics5.4_linux64 <case> <dirout> -gpu ..." invocation, examples/main/01_DamBreak/xCaseDambreak_linux64_GPU.sh:55-60).
// What is printed also goes to <dirout>/Run.out, and the one-line summary to <dirout>/Run.csv, like the reference's JLog2 and
// JSph::ShowResume / SaveRes (src/source/JSph.cpp:3471-3560).
#include "../../../include/b2s_host.h"
#include <cstdarg>
#include <cstdio>
#include <ctime>
#include <string>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <cerrno>
#include <signal.h>
#include <unistd.h>

static std::string g_startfiles, g_tag; static int g_world=0;   // NCCL id file and roll-call files of a multi-process run
static void remove_startfiles(){
  if(g_startfiles.empty())return;
  // every rank has marked its answer "in" once it had read the others' (with a collective initialisation behind it that is long ago;
  // the wait only matters for a transport that connects without one, like the stand-in device of the tests)
  for(int tries=0;tries<100;tries++){
    bool all=true;
    for(int k=1;k<g_world && all;k++){
      char st[8]="", tg[48]=""; FILE* g=fopen((g_startfiles+".r"+std::to_string(k)).c_str(),"r");
      if(g){ if(fscanf(g,"%7s %47s",st,tg)!=2){ st[0]=0; tg[0]=0; } fclose(g); if(!strcmp(tg,g_tag.c_str()) && strcmp(st,"in"))all=false; }
    }
    if(all)break;
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
  }
  remove(g_startfiles.c_str());
  for(int k=0;k<g_world;k++)remove((g_startfiles+".r"+std::to_string(k)).c_str());
}

static FILE* g_runout=NULL;
static void say(const char* fmt,...){
  va_list ap;
  va_start(ap,fmt); vprintf(fmt,ap); va_end(ap);
  if(g_runout){ va_start(ap,fmt); vfprintf(g_runout,fmt,ap); va_end(ap); fflush(g_runout); }
}

int main(int argc,char** argv){
  if(argc<3){
    printf("usage: %s <case (without .xml)> <dirout> [-symplectic|-verlet] [-wendland|-cubic] [-viscoart:v|-viscolam:v|-viscolamsps:v]\n"
           "          [-ddt:0-3] [-ddtvalue:v] [-dbc|-mdbc] [-cellmode:full|half] [-tmax:t] [-tout:t] [-toutx:t] [-nsteps:n] [-svsteps:1] [-sv:none]\n",argv[0]);
    return 1;
  }
  const auto tstart=std::chrono::steady_clock::now();
  std::string opts;
  for(int c=3;c<argc;c++){ opts+=argv[c]; opts+=' '; }
  b2s_sim* sim=NULL;
  // Several GPUs: one process per GPU, started by any launcher that sets RANK / WORLD_SIZE / LOCAL_RANK (torchrun, mpirun wrappers,
  // a shell loop).  Rank 0 makes the NCCL id and leaves it in a file every rank can read (one node); each rank then loads its slab
  // of the case, connects (csrc/slab_comm.cu) and b2s_sim_run steps collectively, writing its piece of every PART.
  const char* ews=getenv("WORLD_SIZE"); const int world=(ews? atoi(ews): 1);
  const int rank=(getenv("RANK")? atoi(getenv("RANK")): 0), local=(getenv("LOCAL_RANK")? atoi(getenv("LOCAL_RANK")): rank);
  int r;
  if(world>1){
    std::string idfile=(getenv("B2S_NCCL_ID_FILE")? getenv("B2S_NCCL_ID_FILE"): std::string("/tmp/b2s_nccl_")+(getenv("MASTER_PORT")? getenv("MASTER_PORT"): "0")+".id");
    // Start-up protocol (one node, files next to each other).  Rank 0 publishes `idfile` = the 128-byte NCCL id followed by a tag
    // "<pid of rank 0>-<time>"; every rank loads its slab and answers a ROLL CALL in `idfile.r<rank>` ("ok <tag>" / "fail <tag>")
    // before anybody enters the communicator's initialisation: a rank that could not load (no such device, bad switch, case not
    // found) stops its peers instead of leaving them inside that collective for ever, and a rank that never shows up is given up
    // after B2S_START_TIMEOUT seconds (default 600).  Left-overs of a run that died cannot be mistaken for this one: answers count
    // only with the tag of the id a rank holds, an id whose rank 0 is not alive is not taken, and a rank that picked up an older id
    // re-reads the file while it waits.
    uint8_t id[128]; char tag[48]="";
    auto read_id=[&](uint8_t* idout,char* tagout)->bool{   // a complete id file of a LIVING rank 0
      unsigned char buf[128+48]; FILE* f=fopen(idfile.c_str(),"rb"); if(!f)return false;
      const size_t n=fread(buf,1,sizeof(buf),f); fclose(f);
      if(n<=130 || n>=sizeof(buf))return false;
      char tg[48]; memcpy(tg,buf+128,n-128); tg[n-128]=0;
      const long pid0=atol(tg);
      if(pid0<=0 || (!getenv("B2S_START_NOPIDCHECK") && kill(pid_t(pid0),0)!=0 && errno==ESRCH))return false;   // (ranks in separate pid namespaces: set B2S_START_NOPIDCHECK)
      memcpy(idout,buf,128); strcpy(tagout,tg);
      return true;
    };
    auto say_rollcall=[&](const char* word){ FILE* f=fopen((idfile+".r"+std::to_string(rank)+".tmp").c_str(),"w"); if(f){ fprintf(f,"%s %s",word,tag); fclose(f); rename((idfile+".r"+std::to_string(rank)+".tmp").c_str(),(idfile+".r"+std::to_string(rank)).c_str()); } };
    if(rank==0){
      if(b2s_slab_comm_id(id)){ fprintf(stderr,"*** NCCL is not available\n"); return 1; }
      snprintf(tag,sizeof(tag),"%ld-%lld",long(getpid()),(long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count());
      for(int k=0;k<world;k++)remove((idfile+".r"+std::to_string(k)).c_str());   // left-overs of a run that died
      FILE* f=fopen((idfile+".tmp").c_str(),"wb"); if(!f || fwrite(id,1,128,f)!=128 || fwrite(tag,1,strlen(tag),f)!=strlen(tag)){ fprintf(stderr,"*** cannot write %s\n",idfile.c_str()); return 1; }
      fclose(f); rename((idfile+".tmp").c_str(),idfile.c_str());
    }
    else{
      bool got=false;
      for(int k=0;k<6000 && !got;k++){   // up to 10 minutes: rank 0 may still be reading a large case
        got=read_id(id,tag);
        if(!got)std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
      if(!got){ fprintf(stderr,"*** rank %d: no NCCL id in %s\n",rank,idfile.c_str()); return 1; }
    }
    r=b2s_sim_load_slab(argv[1],opts.c_str(),local,rank,world,&sim);
    if(!r)r=b2s_sim_set_dirout(sim,argv[2]);
    say_rollcall(r? "fail": "ok");
    if(!r){
      const char* et=getenv("B2S_START_TIMEOUT"); const double tmax=(et? atof(et): 600.0);
      const auto tw=std::chrono::steady_clock::now();
      int bad=-1; bool all=false;
      while(!all && bad<0){
        if(rank!=0){   // rank 0 of THIS run may have started after the id this rank picked up was written
          uint8_t id2[128]; char tag2[48];
          if(read_id(id2,tag2) && strcmp(tag2,tag)){ memcpy(id,id2,128); strcpy(tag,tag2); say_rollcall("ok"); }
        }
        all=true;
        for(int k=0;k<world && bad<0;k++){
          char st[8]="", tg[48]=""; FILE* g=fopen((idfile+".r"+std::to_string(k)).c_str(),"r");
          if(g){ if(fscanf(g,"%7s %47s",st,tg)!=2){ st[0]=0; tg[0]=0; } fclose(g); }
          if(strcmp(tg,tag))all=false;                    // no answer yet, or the answer of another run
          else if(!strcmp(st,"fail"))bad=k;
          else if(strcmp(st,"ok") && strcmp(st,"in"))all=false;
        }
        if(!all && bad<0){
          if(std::chrono::duration<double>(std::chrono::steady_clock::now()-tw).count()>tmax){ bad=world; break; }
          std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
      }
      if(bad>=0){
        if(bad<world)fprintf(stderr,"*** rank %d: rank %d could not start, stopping\n",rank,bad);
        else fprintf(stderr,"*** rank %d: not every rank of %d came up within %g s (B2S_START_TIMEOUT), stopping\n",rank,world,tmax);
        if(rank==0){ remove(idfile.c_str()); remove((idfile+".r0").c_str()); }   // nothing of rank 0 stays for a later run to find
        return 1;
      }
    }
    if(!r){ say_rollcall("in"); r=b2s_sim_slab_connect(sim,id); }   // "in": this rank has seen every answer (rank 0 removes the files only after that)
    g_startfiles=idfile; g_world=world; g_tag=tag;   // removed by rank 0 when its run ends (every rank has long read them by then)
  }
  else r=b2s_sim_load(argv[1],argv[2],opts.c_str(),-1,0,&sim);
  if(r){ fprintf(stderr,"*** load failed (%d): %s\n",r,b2s_host_last_error()); return 1; }
  if(rank!=0){   // the other ranks run silently; rank 0 reports
    double t=0; uint32_t steps=0;
    r=b2s_sim_run(sim,0,&t,&steps);
    if(r)fprintf(stderr,"*** rank %d: run failed (%d): %s\n",rank,r,b2s_host_last_error());
    b2s_sim_free(sim);
    return r? 1: 0;
  }
  std::string dirout=argv[2];
  if(!dirout.empty() && dirout.back()!='/')dirout+='/';
  if(!dirout.empty())g_runout=fopen((dirout+"Run.out").c_str(),"w");   // the directory exists: b2s_sim_load created it
  char date[64]="";
  { const time_t now=time(NULL); struct tm tmv; if(localtime_r(&now,&tmv))strftime(date,sizeof(date),"%d-%m-%Y %H:%M:%S",&tmv); }
  b2s_params p; b2s_sim_params(sim,&p);
  say("b200sph v0.1 (DualSPHysics v5.4 case interface)  %s\n",date);
  say("Case: %s   Options: %s\n",argv[1],opts.c_str());
  say("b200sph: np=%u (fixed %u, moving %u, floating %u, fluid %u)  h=%g  cells=%ux%ux%u\n",p.casenp,p.casenfixed,p.casenmoving,p.casenfloat,p.casenfluid,
    (double)p.kernelh,p.map_cells[0],p.map_cells[1],p.map_cells[2]);
  const auto t0=std::chrono::steady_clock::now();
  double t=0; uint32_t steps=0;
  r=b2s_sim_run(sim,0,&t,&steps);
  const double secs=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  const double ttot=std::chrono::duration<double>(std::chrono::steady_clock::now()-tstart).count();
  if(r){
    fprintf(stderr,"*** run failed (%d): %s\n",r,b2s_host_last_error());
    if(g_runout){ fprintf(g_runout,"*** run failed (%d): %s\n",r,b2s_host_last_error()); fclose(g_runout); }
    b2s_sim_free(sim); remove_startfiles(); return 1;
  }
  say("b200sph: %u steps, simulated time %g s, runtime %.3f s, %.4g particle-steps/s\n",steps,t,secs,secs>0? double(p.casenp)*steps/secs: 0.0);
  uint32_t parts=0, nout[4]={0,0,0,0};
  b2s_sim_stats(sim,&parts,nout);
  b2s_info inf; inf.np=0; inf.nct=0;
  { b2s_ctx* ctx=NULL; if(b2s_sim_ctx(sim,&ctx)==0 && ctx)b2s_get_info(ctx,&inf); }
  //-Resume of the execution with the reference's labels (JSph::ShowResume).
  say("\n[Simulation finished  %s]\n",date);
  say("Particles of simulation (initial): %u\n",p.casenp);
  say("Excluded particles...............: %u\n",nout[0]);
  if(nout[2])say("Excluded particles due to Density: %u\n",nout[2]);
  if(nout[3])say("Excluded particles due to Velocity: %u\n",nout[3]);
  say("Total Runtime....................: %f sec.\n",ttot);
  say("Simulation Runtime...............: %f sec.\n",secs);
  say("Runtime per physical second......: %f sec.\n",t>0? secs/t: 0.0);
  say("Steps per second.................: %f\n",secs>0? steps/secs: 0.0);
  say("Steps of simulation..............: %u\n",steps);
  say("PART files.......................: %u\n",parts);
  say("Maximum number of particles......: %u\n",inf.np);
  say("Maximum number of cells..........: %u\n",inf.nct);
  if(!dirout.empty()){   // Run.csv: the reference's columns (JSph::GetResInfo); the ones without a counterpart here stay empty
    FILE* f=fopen((dirout+"Run.csv").c_str(),"w");
    if(f){
      fprintf(f,"#RunName;Rcode-VersionInfo;DateTime;Np;TSimul;TSeg;TTotal;MemCpu;MemGpu;MemGpuCells;Steps;GPIPS;PhysicalTime;PartFiles;PartsOut;"
                "MaxParticles;MaxCells;Hardware;RunMode;Configuration;Nbound;Nfixed;Dp;H;PartsOutRho;PartsOutVel\n");
      std::string name=argv[1]; { const size_t k=name.find_last_of('/'); if(k!=std::string::npos)name=name.substr(k+1); }
      fprintf(f,"%s;b200sph v0.1;%s;%u;%g;%g;%g;;;;%u;;%.15g;%u;%u;%u;%u;B200;%s;%s;%u;%u;%g;%g;%u;%u\n",name.c_str(),date,p.casenp,
        secs,t>0? secs/t: 0.0,ttot,steps,t,parts,nout[0],inf.np,inf.nct,world>1? "MultiGPU-slabs": "SingleGPU",opts.c_str(),p.casenp-p.casenfluid,p.casenfixed,(double)p.dp,(double)p.kernelh,nout[2],nout[3]);
      fclose(f);
    }
  }
  b2s_sim_free(sim);
  remove_startfiles();
  say("\nFinished execution (code=0).\n");
  if(g_runout)fclose(g_runout);
  return 0;
}
