// motionref.h -- PartMotionRef.ibi4: three reference particles per moving / floating body and their positions at every PART, the
// file the reference's post-processing tools use to rebuild the motion of the bodies (JDsPartMotionSave.cpp:118-235,
// JComputeMotionRef.cpp:127-236, JPartMotRefBi4Save.cpp:149-292).  Host code of the case interface; no device work.
#ifndef B2S_HOST_MOTIONREF_H_
#define B2S_HOST_MOTIONREF_H_
#include <cstdint>
#include <string>
#include <vector>
namespace b2s {

struct MotRefBody {
  uint16_t mkbound; uint32_t begin, np;      // block of ids [begin,begin+np)
  uint32_t nid; uint32_t id[3]; double ps[3][3]; double dis[3];
};

class MotionRefSave {
public:
  // Bodies in id order: moving blocks first, then floating ones.
  void add_body(bool floating,uint16_t mkbound,uint32_t begin,uint32_t np);
  // Picks the reference particles from the initial state.  idp[] ascending (the particles of the bodies), pos[3n].
  void config(uint32_t n,const uint32_t* idp,const double* pos);
  // Appends PART_<cpart> (head written when `fresh`, JPartMotRefBi4Save::SaveInitial + SavePart).
  void save_part(const std::string& dirdata,const std::string& appname,double timeout,unsigned mkboundfirst,bool fresh,
                 unsigned cpart,double timestep,unsigned step,uint32_t n,const uint32_t* idp,const double* pos);
  // The extra-frequency file PartMotionRef2.ibi4 (TimeOutExtra / -toutx; JDsPartMotionSave::AddDataExtra*, JPartMotRefBi4Save::
  // AddDataPart / SaveStoredData, JPartMotRefBi4Save.cpp:225-278): `count` records (time, step, 3 positions per body) collected while
  // PART <cpart> was the next one to be written; one record is stored like a PART of the main file, several as arrays.
  void save_extra(const std::string& dirdata,const std::string& appname,double timeout2,unsigned mkboundfirst,bool fresh,
                  unsigned cpart,unsigned count,const double* times,const uint32_t* steps,const double* posref);
  // ids of the reference particles, three per body in file order (what posref of save_extra holds, 9 doubles per body)
  std::vector<uint32_t> ref_ids()const;
  size_t body_count()const{ return bodies.size(); }
  const MotRefBody& body(size_t c)const{ return bodies[c]; }
private:
  std::vector<MotRefBody> bodies; unsigned nmoving=0,nfloat=0;
};

} // namespace b2s
#endif
