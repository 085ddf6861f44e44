"""The command line's log and summary files (Run.out, Run.csv; reference: JSph::ShowResume / GetResInfo, JSph.cpp:3471-3560) without a
GPU: host/main.cpp is linked against a stand-in for libb200sph.so whose b2s_sim_* entry points return canned values, so everything
the CLI does around the library -- option string, log tee, resume labels, CSV columns, exit code -- runs here."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STUB = r'''
#include "b2s_host.h"
#include <cstring>
#include <sys/stat.h>
struct b2s_sim{int x;}; struct b2s_ctx{int y;};
static b2s_sim S; static b2s_ctx C;
extern "C" {
int b2s_sim_load(const char* c,const char* dirout,const char*,int,uint32_t,b2s_sim** out){ if(strstr(c,"missing"))return -8; mkdir(dirout,0777); *out=&S; return 0; }
int b2s_sim_params(b2s_sim*,b2s_params* p){ memset(p,0,sizeof(*p)); p->casenp=1000; p->casenfluid=800; p->casenfixed=150; p->casenmoving=50; p->kernelh=0.01f; p->dp=0.005f; return 0; }
int b2s_sim_run(b2s_sim*,uint32_t,double* t,uint32_t* st){ *t=0.02; *st=24; return 0; }
int b2s_sim_stats(b2s_sim*,uint32_t* parts,uint32_t* e){ *parts=5; e[0]=3; e[1]=0; e[2]=2; e[3]=1; return 0; }
int b2s_sim_ctx(b2s_sim*,b2s_ctx** c){ *c=&C; return 0; }
int b2s_get_info(b2s_ctx*,b2s_info* i){ memset(i,0,sizeof(*i)); i->np=997; i->nct=4321; return 0; }
int b2s_sim_free(b2s_sim*){ return 0; }
int b2s_slab_comm_id(uint8_t*){ return -1; }                       // the multi-GPU branch of main.cpp is not taken here (WORLD_SIZE unset)
int b2s_sim_load_slab(const char*,const char*,int,int,int,b2s_sim**){ return -1; }
int b2s_sim_set_dirout(b2s_sim*,const char*){ return -1; }
int b2s_sim_slab_connect(b2s_sim*,const uint8_t*){ return -1; }
const char* b2s_host_last_error(void){ return "xml: cannot open"; }
}
'''


def test_cli_writes_run_out_and_run_csv(tmp_path):
    (tmp_path / "stub.cpp").write_text(STUB)
    inc = os.path.join(ROOT, "include")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I", inc, "stub.cpp", "-o", "libb200sph.so"], cwd=tmp_path, check=True)
    subprocess.run(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "dualsphysics_b200", "csrc", "host", "main.cpp"), "-o", "cli",
                    "-L.", "-lb200sph", "-Wl,-rpath,$ORIGIN"], cwd=tmp_path, check=True)
    out = tmp_path / "out"
    r = subprocess.run([str(tmp_path / "cli"), "dir/Case", str(out), "-symplectic", "-tmax:0.02"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.rstrip().endswith("Finished execution (code=0).")
    log = (out / "Run.out").read_text()
    assert log == r.stdout                                           # the log file is what was printed
    for needle in ("Particles of simulation (initial): 1000", "Excluded particles...............: 3", "Excluded particles due to Density: 2",
                   "Steps of simulation..............: 24", "PART files.......................: 5", "Maximum number of cells..........: 4321"):
        assert needle in log, needle
    head, row = (out / "Run.csv").read_text().strip().splitlines()
    h, v = head.split(";"), row.split(";")
    assert len(h) == len(v) == 26 and h[0] == "#RunName"
    rec = dict(zip(h, v))
    assert rec["#RunName"] == "Case" and rec["Np"] == "1000" and rec["Steps"] == "24" and rec["PartFiles"] == "5" and rec["PartsOut"] == "3"
    assert rec["Nbound"] == "200" and rec["Nfixed"] == "150" and rec["PartsOutRho"] == "2" and rec["PartsOutVel"] == "1"
    assert rec["Configuration"].split() == ["-symplectic", "-tmax:0.02"] and float(rec["PhysicalTime"]) == 0.02
    # a failing load: message on stderr, exit code 1, nothing written
    r = subprocess.run([str(tmp_path / "cli"), "missing/Case", str(tmp_path / "out2")], capture_output=True, text=True)
    assert r.returncode == 1 and "load failed (-8)" in r.stderr and not (tmp_path / "out2").exists()
    # usage
    r = subprocess.run([str(tmp_path / "cli")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("usage:")
