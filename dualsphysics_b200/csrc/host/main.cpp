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
    uint8_t id[128];
    if(rank==0){
      if(b2s_slab_comm_id(id)){ fprintf(stderr,"*** NCCL is not available\n"); return 1; }
      FILE* f=fopen((idfile+".tmp").c_str(),"wb"); if(!f || fwrite(id,1,128,f)!=128){ fprintf(stderr,"*** cannot write %s\n",idfile.c_str()); return 1; }
      fclose(f); rename((idfile+".tmp").c_str(),idfile.c_str());
    }
    else{
      bool got=false;
      for(int k=0;k<6000 && !got;k++){   // up to 10 minutes: rank 0 may still be reading a large case
        FILE* f=fopen(idfile.c_str(),"rb");
        if(f){ got=(fread(id,1,128,f)==128); fclose(f); }
        if(!got)std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
      if(!got){ fprintf(stderr,"*** rank %d: no NCCL id in %s\n",rank,idfile.c_str()); return 1; }
    }
    r=b2s_sim_load_slab(argv[1],opts.c_str(),local,rank,world,&sim);
    if(!r)r=b2s_sim_set_dirout(sim,argv[2]);
    if(!r)r=b2s_sim_slab_connect(sim,id);
    if(rank==0)remove(idfile.c_str());   // every rank is connected once rank 0's collective init returns
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
    b2s_sim_free(sim); return 1;
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
  say("\nFinished execution (code=0).\n");
  if(g_runout)fclose(g_runout);
  return 0;
}
