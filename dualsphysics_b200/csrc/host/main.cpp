// main.cpp -- b200sph command line: b200sph <case> <dirout> [options]   (same shape as the reference's
// "DualSPHysics5.4_linux64 <case> <dirout> -gpu ..." invocation, examples/main/01_DamBreak/xCaseDambreak_linux64_GPU.sh:55-60).
#include "../../../include/b2s_host.h"
#include <cstdio>
#include <string>
#include <chrono>

int main(int argc,char** argv){
  if(argc<3){
    printf("usage: %s <case (without .xml)> <dirout> [-symplectic|-verlet] [-wendland|-cubic] [-viscoart:v|-viscolam:v|-viscolamsps:v]\n"
           "          [-ddt:0-3] [-ddtvalue:v] [-dbc|-mdbc] [-cellmode:full|half] [-tmax:t] [-tout:t] [-nsteps:n] [-svsteps:1] [-sv:none]\n",argv[0]);
    return 1;
  }
  std::string opts;
  for(int c=3;c<argc;c++){ opts+=argv[c]; opts+=' '; }
  b2s_sim* sim=NULL;
  int r=b2s_sim_load(argv[1],argv[2],opts.c_str(),-1,0,&sim);
  if(r){ fprintf(stderr,"*** load failed (%d): %s\n",r,b2s_host_last_error()); return 1; }
  b2s_params p; b2s_sim_params(sim,&p);
  printf("b200sph: np=%u (fixed %u, moving %u, floating %u, fluid %u)  h=%g  cells=%ux%ux%u\n",p.casenp,p.casenfixed,p.casenmoving,p.casenfloat,p.casenfluid,
    (double)p.kernelh,p.map_cells[0],p.map_cells[1],p.map_cells[2]);
  const auto t0=std::chrono::steady_clock::now();
  double t=0; uint32_t steps=0;
  r=b2s_sim_run(sim,0,&t,&steps);
  const double secs=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  if(r){ fprintf(stderr,"*** run failed (%d): %s\n",r,b2s_host_last_error()); b2s_sim_free(sim); return 1; }
  printf("b200sph: %u steps, simulated time %g s, runtime %.3f s, %.4g particle-steps/s\n",steps,t,secs,secs>0? double(p.casenp)*steps/secs: 0.0);
  b2s_sim_free(sim);
  printf("\nFinished execution (code=0).\n");
  return 0;
}
