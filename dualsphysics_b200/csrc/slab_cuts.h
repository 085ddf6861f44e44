// slab_cuts.h -- cut planes of the slab decomposition along x (shared by the host driver and the device-side re-cut).
// No reference counterpart: DualSPHysics v5.4 is single-GPU (src/source/DualSphDef.h:96-101).
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>

// Work of one particle in units of a boundary particle: a fluid particle costs two neighbour walks over ~170 pairs per step
// (1.0 us each on the B200 at 10M particles) plus the per-particle passes (0.3 us), a boundary particle one short walk over the
// fluid near it (0.2 us) plus the same passes: about 3:1 (profiles/README.md, launch list of the bench step).
#define B2S_SLAB_WEIGHT_FLUID 3u
#define B2S_SLAB_WEIGHT_BOUND 1u

// Balanced cut planes: prefix sum of the WORK per x cell plane, every cut at the plane whose cumulative work is nearest to
// r*total/G; a slab is at least minw planes wide (the halo only comes from the direct neighbours).  Same rule as
// dualsphysics_b200/slab.py compute_cuts_hist (tests compare both).
static inline std::vector<uint32_t> b2s_slab_compute_cuts(const std::vector<uint64_t>& hist,int nranks,uint32_t minw){
  const uint32_t ncx=uint32_t(hist.size());
  std::vector<uint64_t> cum(ncx); uint64_t t=0; for(uint32_t i=0;i<ncx;i++){ t+=hist[i]; cum[i]=t; }
  std::vector<uint32_t> cuts; cuts.push_back(0);
  for(int r=1;r<nranks;r++){
    const double target=double(t)*r/nranks;
    uint32_t i=0; while(i<ncx && double(cum[i])<target)i++;      // searchsorted(cum,target,'left')
    uint32_t c=i+1;
    if(c>=2 && c-1<ncx && fabs(double(cum[c-2])-target)<fabs(double(cum[c-1])-target))c--;
    if(c<cuts.back()+minw)c=cuts.back()+minw;
    if(c>ncx-uint32_t(nranks-r)*minw)c=ncx-uint32_t(nranks-r)*minw;
    cuts.push_back(c);
  }
  cuts.push_back(ncx);
  return cuts;
}
